"""Timing of the tensor-core value path on deeper / wider eligible nets vs the CUDA-core kernels (development aid)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'scripts'))
from gpu_probe import mlp_desc, timeit
from quinn_b200 import ops
for d, hls, N, K in ((3, (64, 64, 64), 10000, 2368), (10, (128, 128), 10000, 1184), (2, (32, 32), 1000, 8192), (3, (64, 64, 64, 64), 10000, 2368)):
    rs = np.random.RandomState(0)
    desc = mlp_desc(d, 1, hls)
    x = rs.rand(N, d) * 2 - 1
    y = np.sin(x.sum(1, keepdims=True))
    for notc in ('1', '0'):
        os.environ['QB_NO_TC'] = notc
        prob = ops.Problem(desc, x, y, 0.05, dtype=torch.float32)
        th = prob.theta(0.2 * rs.randn(K, desc.n_params))
        lp = torch.empty(K, dtype=torch.float64, device='cuda')
        med, best = timeit(lambda: ops.logpost(prob, th, out=lp), reps=5, warm=2)
        print(d, hls, 'QB_NO_TC=' + notc, prob.plan_info(K)['tensor_core'], 'ms %.3f' % med, 'evals/s %.0f' % (K / med * 1e3),
              'TFLOP/s %.1f' % (K * 2.0 * N * desc.macs_per_point() / med / 1e9), flush=True)

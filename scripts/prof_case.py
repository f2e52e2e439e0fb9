"""Run one kernel case a few times (development aid for ncu captures):
   python scripts/prof_case.py c5 grad 2368 [f32|f64]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'scripts'))
from gpu_probe import mlp_desc, timeit
from quinn_b200 import ops
case, mode, K = sys.argv[1], sys.argv[2], int(sys.argv[3])
dt = torch.float64 if len(sys.argv) > 4 and sys.argv[4] == 'f64' else torch.float32
d, hls, N = {'c5': (3, (64, 64), 10000), 'c2': (2, (32, 32), 1000), 'c3': (10, (128, 128), 10000)}[case]
rs = np.random.RandomState(0)
desc = mlp_desc(d, 1, hls)
x = rs.rand(N, d) * 2 - 1
y = np.sin(x.sum(1, keepdims=True))
prob = ops.Problem(desc, x, y, 0.05, dtype=dt)
th = prob.theta(0.2 * rs.randn(K, desc.n_params))
lp = torch.empty(K, dtype=torch.float64, device='cuda'); g = torch.empty_like(th)
fn = (lambda: ops.logpost_grad(prob, th, lp, g)) if mode == 'grad' else (lambda: ops.logpost(prob, th, out=lp))
med, best = timeit(fn, reps=3, warm=2)
S = desc.macs_per_point()
fl = (6.0 * N * S - 2.0 * N * desc.layers[0].n_in * desc.layers[0].n_out) if mode == 'grad' else 2.0 * N * S
print(case, mode, K, prob.plan_info(K, mode == 'grad'), 'ms', med, 'TFLOP/s', K * fl / med / 1e9)

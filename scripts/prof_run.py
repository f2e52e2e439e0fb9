"""One or two launches of a named hot kernel for ncu (development aid).
   python scripts/prof_run.py grad|hmc|amcmc|predict [K]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
from gpu_probe import mlp_desc          # noqa: E402
from quinn_b200 import ops              # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'grad'
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
rs = np.random.RandomState(0)
d, hls, N = 3, (64, 64), 10000
desc = mlp_desc(d, 1, hls)
x = rs.rand(N, d) * 2 * np.pi - np.pi
y = np.sin(x).sum(1, keepdims=True) + 0.05 * rs.randn(N, 1)
prob = ops.Problem(desc, x, y, 0.05, dtype=torch.float32)
th = prob.theta(rs.rand(K, desc.n_params))
if what == 'grad':
    lp = torch.empty(K, dtype=torch.float64, device='cuda')
    g = torch.empty_like(th)
    for _ in range(2):
        ops.logpost_grad(prob, th, lp, g)
elif what == 'hmc':
    st = ops.ChainState(prob, th)
    hm = ops.HmcState(st, epsilon=2e-6, L=3)
    ops.hmc_run(st, hm, 1, None, seed=1)
    ops.hmc_run(st, hm, 2, None, seed=1)
elif what == 'amcmc':
    st = ops.ChainState(prob, th)
    am = ops.AmcmcState(st, gamma=0.01, adapt='diag')
    ops.amcmc_run(st, am, 2, None, seed=1)
    ops.amcmc_run(st, am, 4, None, seed=1)
torch.cuda.synchronize()
print('done', what, K)

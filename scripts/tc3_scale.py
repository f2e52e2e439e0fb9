"""Occupancy / latency experiment for kernel 1 at the config-5 shape: time of ONE wave with 1 and 2 blocks per SM and
the fixed (staging) cost per evaluation from two data sizes.   python scripts/tc3_scale.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
from gpu_probe import mlp_desc, timeit          # noqa: E402
from quinn_b200 import ops                      # noqa: E402

os.environ['QB_SPLIT'] = '1'
rs = np.random.RandomState(0)
desc = mlp_desc(3, 1, (64, 64))
for N in (10112, 5056, 128):
    x = rs.rand(N, 3) * 2 - 1
    y = np.sin(x.sum(1, keepdims=True))
    prob = ops.Problem(desc, x, y, 0.05, dtype=torch.float32)
    for K in (74, 148, 296, 592, 2368):
        th = prob.theta(0.2 * rs.randn(K, desc.n_params))
        lp = torch.empty(K, dtype=torch.float64, device='cuda')
        med, best = timeit(lambda: ops.logpost(prob, th, lp), reps=9, warm=3)
        print(f'N={N} tiles={N // 128} K={K}: best {best * 1e3:.1f} us  median {med * 1e3:.1f} us  -> cycles/tile/block {best * 1e-3 * 1.965e9 / (N / 128) / max(1, K / 296):.0f}', flush=True)

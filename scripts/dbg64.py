"""Development check of the 64-wide fp16-split gradient kernel with one and two blocks per SM (the hand-over race of
profiles/README.md showed up here as non-finite gradient rows in the second resident block).
   python scripts/dbg64.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
from gpu_probe import mlp_desc
from quinn_b200 import ops
def run(K, N, split=None, d=3):
    if split: os.environ['QB_SPLIT'] = str(split)
    elif 'QB_SPLIT' in os.environ: del os.environ['QB_SPLIT']
    rs = np.random.RandomState(0)
    desc = mlp_desc(d, 1, (64, 64))
    x = rs.rand(N, d) * 2 - 1
    y = np.sin(x.sum(1, keepdims=True)) + 0.1 * rs.randn(N, 1)
    th0 = 0.5 * rs.randn(K, desc.n_params)
    prob = ops.Problem(desc, x, y, 0.1, dtype=torch.float32)
    info = prob.plan_info(K, True)
    for rep in range(3):
        lp, g = ops.logpost_grad(prob, th0)
        torch.cuda.synchronize()
        g = g.cpu().numpy(); lp = lp.cpu().numpy()
        bad = np.where(~np.isfinite(g).all(1))[0]
        badlp = np.where(~np.isfinite(lp))[0]
        print(f'K={K} N={N} split={info["splits"]} tc={info["tensor_core"]} rep {rep}: bad grad rows {len(bad)} {bad[:12]} bad lp {len(badlp)} {badlp[:8]}', flush=True)
        if len(bad):
            k = bad[0]
            nanidx = np.where(~np.isfinite(g[k]))[0]
            print('   row', k, 'nan count', len(nanidx), 'first', nanidx[:10], 'last', nanidx[-5:])
run(5, 1000); run(148, 777, 1); run(296, 777, 1); run(300, 777, 1); run(300, 777); run(600, 256, 1); run(1200, 128, 1); run(2000, 100, 1)

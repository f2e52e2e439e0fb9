"""Development check of the 128-wide tensor-core GRADIENT path (qb_tg8.cuh): fp32 tensor-core kernel vs the fp64
CUDA-core kernel, error reported per parameter block so that a wrong stage is visible at once; then timings.
   python scripts/tg8_check.py [quick]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
from gpu_probe import mlp_desc, timeit          # noqa: E402
from tcg_check import blocks                    # noqa: E402
from quinn_b200 import ops                      # noqa: E402


def run(d, N, K, sigma=0.1, seed=0, prior=False, wscale=0.5, xscale=1.0, yscale=1.0):
    rs = np.random.RandomState(seed)
    desc = mlp_desc(d, 1, (128, 128))
    x = (rs.rand(N, d) * 2 - 1) * xscale
    y = (np.sin(x.sum(1, keepdims=True) / xscale) + 0.1 * rs.randn(N, 1)) * yscale
    th0 = wscale * rs.randn(K, desc.n_params)
    kw = dict(prior_sigma=0.7, prior_anchor=0.1 * rs.randn(desc.n_params)) if prior else {}
    res = {}
    for name, dt, env in (('f64', torch.float64, {}), ('simt', torch.float32, {'QB_NO_TG8': '1'}), ('tc', torch.float32, {'QB_NO_TG8': '0'})):
        os.environ.update(env)
        prob = ops.Problem(desc, x, y, sigma, dtype=dt, **kw)
        info = prob.plan_info(K, True)
        lp, g = ops.logpost_grad(prob, th0)
        torch.cuda.synchronize()
        res[name] = (lp.cpu().numpy(), g.double().cpu().numpy(), info)
    ref_lp, ref_g, _ = res['f64']
    line = (f'net {d}-128-128-1 N={N} K={K} prior={prior} w={wscale} x={xscale} y={yscale} sigma={sigma} '
            f'plan(tc)={res["tc"][2]["tensor_core"]} splits={res["tc"][2]["splits"]}:')
    for name in ('simt', 'tc'):
        lp, g, _ = res[name]
        e_lp = np.max(np.abs(lp - ref_lp) / np.abs(ref_lp))
        gmax = np.abs(ref_g).max(axis=1, keepdims=True)
        parts = []
        for bn, a, b in blocks(desc):
            parts.append(f'{bn} {np.max(np.abs(g[:, a:b] - ref_g[:, a:b]) / gmax):.2e}')
        line += f'\n   {name:5s} lp {e_lp:.2e} | ' + ' '.join(parts)
    print(line, flush=True)


if __name__ == '__main__':
    quick = len(sys.argv) > 1 and sys.argv[1] == 'quick'
    run(10, 100, 2)
    run(10, 128, 3)
    run(10, 129, 1)
    run(10, 1000, 5, prior=True)
    run(10, 777, 200)
    run(3, 300, 4)
    run(7, 300, 4)
    run(11, 300, 4)
    run(1, 50, 3)
    run(10, 500, 4, wscale=0.05, xscale=100.0, yscale=30.0, sigma=2.0)
    run(10, 500, 4, wscale=3.0, xscale=1e-3, yscale=1e-2, sigma=0.01)
    if not quick:
        d, N, K = 10, 100000, 128
        rs = np.random.RandomState(0)
        desc = mlp_desc(d, 1, (128, 128))
        x = rs.rand(N, d) * 2 - 1
        y = np.sin(x.sum(1, keepdims=True))
        S = desc.macs_per_point()
        F_vg = 6.0 * N * S - 2.0 * N * desc.layers[0].n_in * desc.layers[0].n_out
        for no in ('1', '0'):
            os.environ['QB_NO_TG8'] = no
            prob = ops.Problem(desc, x, y, 0.05, dtype=torch.float32)
            th = prob.theta(0.2 * rs.randn(K, desc.n_params))
            lp = torch.empty(K, dtype=torch.float64, device='cuda')
            g = torch.empty_like(th)
            med, best = timeit(lambda: ops.logpost_grad(prob, th, lp, g), reps=5, warm=2)
            print(f'net {d}-128-128 N={N} K={K} QB_NO_TG8={no}: grad ms {med:.3f} evals/s {K / med * 1e3:.4g} TFLOP/s {K * F_vg / med / 1e9:.2f}', flush=True)

"""Development check of the tensor-core GRADIENT path (qb_tcg.cuh): fp32 tensor-core kernel vs the fp64 CUDA-core
kernel, error reported per parameter block so that a wrong stage is visible at once; then timings.
   python scripts/tcg_check.py [quick]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
from gpu_probe import mlp_desc, timeit          # noqa: E402
from quinn_b200 import ops                      # noqa: E402


def blocks(desc):
    out = []
    for l, L in enumerate(desc.layers):
        out.append((f'W{l}', L.w_off, L.w_off + L.n_in * L.n_out))
        if L.b_off >= 0:
            out.append((f'b{l}', L.b_off, L.b_off + L.n_out))
    return out


def run(d, hls, N, K, sigma=0.1, seed=0, prior=False):
    rs = np.random.RandomState(seed)
    desc = mlp_desc(d, 1, hls)
    x = rs.rand(N, d) * 2 - 1
    y = np.sin(x.sum(1, keepdims=True)) + 0.1 * rs.randn(N, 1)
    th0 = 0.5 * rs.randn(K, desc.n_params)
    kw = dict(prior_sigma=0.7, prior_anchor=0.1 * rs.randn(desc.n_params)) if prior else {}
    res = {}
    for name, dt, env in (('f64', torch.float64, {}), ('simt', torch.float32, {'QB_NO_TCG': '1'}), ('tc', torch.float32, {'QB_NO_TCG': '0'})):
        os.environ.update(env)
        prob = ops.Problem(desc, x, y, sigma, dtype=dt, **kw)
        info = prob.plan_info(K, True)
        lp, g = ops.logpost_grad(prob, th0)
        torch.cuda.synchronize()
        res[name] = (lp.cpu().numpy(), g.double().cpu().numpy(), info)
    ref_lp, ref_g, _ = res['f64']
    line = f'net {d}-{hls}-1 N={N} K={K} prior={prior} plan(tc)={res["tc"][2]["tensor_core"]} splits={res["tc"][2]["splits"]}:'
    for name in ('simt', 'tc'):
        lp, g, _ = res[name]
        e_lp = np.max(np.abs(lp - ref_lp) / np.abs(ref_lp))
        gmax = np.abs(ref_g).max(axis=1, keepdims=True)
        parts = []
        for bn, a, b in blocks(desc):
            parts.append(f'{bn} {np.max(np.abs(g[:, a:b] - ref_g[:, a:b]) / gmax):.2e}')
        line += f'\n   {name:5s} lp {e_lp:.2e} | ' + ' '.join(parts)
    print(line, flush=True)
    lp, g, _ = res['tc']
    return np.max(np.abs(g - ref_g) / np.abs(ref_g).max(axis=1, keepdims=True))


if __name__ == '__main__':
    quick = len(sys.argv) > 1 and sys.argv[1] == 'quick'
    run(3, (64, 64), 100, 2)
    run(3, (64, 64), 128, 3)
    run(3, (64, 64), 129, 1)
    run(3, (64, 64), 1000, 5, prior=True)
    run(3, (64, 64), 777, 300)
    run(2, (32, 32), 1000, 50)
    run(2, (32, 32), 50, 3)
    run(6, (64, 64), 300, 4)
    run(7, (32, 32), 300, 4)
    if not quick:
        for d, hls, N, K in ((3, (64, 64), 10000, 2368), (2, (32, 32), 1000, 4096)):
            rs = np.random.RandomState(0)
            desc = mlp_desc(d, 1, hls)
            x = rs.rand(N, d) * 2 - 1
            y = np.sin(x.sum(1, keepdims=True))
            S = desc.macs_per_point()
            F_vg = 6.0 * N * S - 2.0 * N * desc.layers[0].n_in * desc.layers[0].n_out
            for notcg in ('1', '0'):
                os.environ['QB_NO_TCG'] = notcg
                prob = ops.Problem(desc, x, y, 0.05, dtype=torch.float32)
                th = prob.theta(0.2 * rs.randn(K, desc.n_params))
                lp = torch.empty(K, dtype=torch.float64, device='cuda')
                g = torch.empty_like(th)
                med, best = timeit(lambda: ops.logpost_grad(prob, th, lp, g), reps=5, warm=2)
                print(f'net {d}-{hls} N={N} K={K} QB_NO_TCG={notcg}: grad ms {med:.3f} evals/s {K / med * 1e3:.4g} TFLOP/s {K * F_vg / med / 1e9:.2f}', flush=True)

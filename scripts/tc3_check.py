"""Development check of the warp-specialised tensor-core VALUE path (qb_tc3.cuh): fp32 log-posterior of the hot shape
(in<=7 -> 64 -> 64 -> 1, tanh) against the fp64 CUDA-core kernel, next to the previous tensor-core loop (QB_NO_V3=1);
then timings of kernel 1 at the config-5 shape.
   python scripts/tc3_check.py [quick]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
from gpu_probe import mlp_desc, timeit          # noqa: E402
from quinn_b200 import ops                      # noqa: E402


def run(d, N, K, sigma=0.1, seed=0, wscale=0.5, prior=False, H=64):
    rs = np.random.RandomState(seed)
    desc = mlp_desc(d, 1, (H, H))
    x = rs.rand(N, d) * 2 - 1
    y = np.sin(x.sum(1, keepdims=True)) + 0.1 * rs.randn(N, 1)
    th0 = wscale * rs.randn(K, desc.n_params)
    kw = dict(prior_sigma=0.7, prior_anchor=0.1 * rs.randn(desc.n_params)) if prior else {}
    res = {}
    for name, dt, env in (('f64', torch.float64, {}), ('old', torch.float32, {'QB_NO_V3': '1'}), ('v3', torch.float32, {'QB_NO_V3': '0'})):
        os.environ.update(env)
        prob = ops.Problem(desc, x, y, sigma, dtype=dt, **kw)
        info = prob.plan_info(K, False)
        lp = ops.logpost(prob, th0)
        torch.cuda.synchronize()
        res[name] = (lp.cpu().numpy(), info)
    ref = res['f64'][0]
    line = f'net {d}-{H}-{H}-1 N={N} K={K} w~{wscale} prior={prior} plan(v3)={res["v3"][1]["tensor_core"]} threads={res["v3"][1]["threads"]} splits={res["v3"][1]["splits"]}:'
    for name in ('old', 'v3'):
        e = np.abs(res[name][0] - ref) / np.abs(ref)
        line += f'  {name} max {e.max():.2e} mean {e.mean():.2e}'
    print(line, flush=True)


if __name__ == '__main__':
    quick = len(sys.argv) > 1 and sys.argv[1] == 'quick'
    run(3, 100, 2)
    run(3, 128, 3)
    run(3, 129, 1)
    run(3, 1000, 5, prior=True)
    run(3, 777, 300)
    run(3, 10000, 64)
    run(3, 10000, 64, wscale=3.0)
    run(3, 10000, 64, wscale=0.02)
    run(1, 300, 4)
    run(2, 300, 4)
    run(5, 300, 4)
    run(7, 1000, 4)
    for d, N, K in ((10, 100, 2), (10, 129, 3), (10, 1000, 5), (10, 10000, 16), (1, 300, 4), (15, 777, 4), (8, 500, 3)):
        run(d, N, K, H=128, wscale=0.3)
    run(10, 5000, 8, H=128, wscale=2.0)
    # kernel 4 (predict): 64- and 128-wide, against the fp64 CUDA-core kernel
    for d, H, M, N in ((3, 64, 5, 1000), (10, 128, 7, 3333), (10, 128, 3, 129)):
        rs = np.random.RandomState(3)
        desc = mlp_desc(d, 1, (H, H))
        xx = rs.rand(N, d) * 2 - 1
        th = 0.4 * rs.randn(M, desc.n_params)
        ref = ops.predict(desc, th, xx, dtype=torch.float64, want_out=True, want_moments=False)[0].cpu().numpy()
        outs = {}
        for name, env in (('old', '1'), ('v3', '0')):
            os.environ['QB_NO_V3'] = env
            outs[name] = ops.predict(desc, th, xx, dtype=torch.float32, want_out=True, want_moments=False)[0].double().cpu().numpy()
        sc = np.abs(ref).max()
        print(f'predict {d}-{H}-{H}-1 M={M} N={N}: old max err {np.abs(outs["old"] - ref).max() / sc:.2e}  v3 max err {np.abs(outs["v3"] - ref).max() / sc:.2e}', flush=True)
    if not quick:
        d, N, K = 3, 10000, 2368
        rs = np.random.RandomState(0)
        desc = mlp_desc(d, 1, (64, 64))
        x = rs.rand(N, d) * 2 - 1
        y = np.sin(x.sum(1, keepdims=True))
        F_v = 2.0 * N * desc.macs_per_point()
        for nov3 in ('1', '0'):
            os.environ['QB_NO_V3'] = nov3
            prob = ops.Problem(desc, x, y, 0.05, dtype=torch.float32)
            th = prob.theta(0.2 * rs.randn(K, desc.n_params))
            lp = torch.empty(K, dtype=torch.float64, device='cuda')
            med, best = timeit(lambda: ops.logpost(prob, th, lp), reps=7, warm=2)
            print(f'net 3-64-64-1 N={N} K={K} QB_NO_V3={nov3}: logpost ms {med:.3f} evals/s {K / med * 1e3:.4g} TFLOP/s {K * F_v / med / 1e9:.2f}', flush=True)
        # config-3 predictive: 256 members x 1e6 points, 10-128-128-1
        d, H, M, N = 10, 128, 256, 1000000
        desc = mlp_desc(d, 1, (H, H))
        xx = torch.as_tensor(rs.rand(N, d), dtype=torch.float32, device='cuda')
        th = torch.as_tensor((2 * rs.rand(M, desc.n_params) - 1) / np.sqrt(H), dtype=torch.float32, device='cuda')
        cnet = desc.to_c()
        for nov3 in ('1', '0'):
            os.environ['QB_NO_V3'] = nov3
            med, best = timeit(lambda: ops.predict(desc, th, xx, dtype=torch.float32, want_out=False, want_moments=True, cnet=cnet), reps=5, warm=2)
            print(f'predict 10-128-128-1 M={M} N={N} QB_NO_V3={nov3}: ms {med:.2f} member-points/s {M * N / med * 1e3:.4g}', flush=True)

"""Development check of the tensor-core value path: TC vs SIMT (QB_NO_TC=1) vs an fp64 evaluation, and timing.
   python scripts/tc_check.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'scripts'))
from gpu_probe import mlp_desc, timeit
from quinn_b200 import ops


def run(d, hls, o, N, K, seed=0):
    rs = np.random.RandomState(seed)
    desc = mlp_desc(d, o, hls)
    x = rs.rand(N, d) * 2 - 1
    y = np.sin(x.sum(1, keepdims=True)) + 0.1 * rs.randn(N, o)
    th0 = 0.5 * rs.randn(K, desc.n_params)
    out = {}
    for name, dt, notc in (('f64', torch.float64, '1'), ('simt', torch.float32, '1'), ('tc', torch.float32, '0')):
        os.environ['QB_NO_TC'] = notc
        prob = ops.Problem(desc, x, y, 0.1, dtype=dt)
        th = prob.theta(th0)
        out[name] = ops.logpost(prob, th).cpu().numpy()
    ref = out['f64']
    e_simt = np.max(np.abs(out['simt'] - ref) / np.abs(ref)); e_tc = np.max(np.abs(out['tc'] - ref) / np.abs(ref))
    print(f'net {d}-{hls}-{o} N={N} K={K}: rel err simt {e_simt:.3g}  tc {e_tc:.3g}   lp[0] {ref[0]:.6f} {out["tc"][0]:.6f}', flush=True)
    return e_tc


if __name__ == '__main__':
    run(3, (64, 64), 1, 1000, 8)
    run(3, (64, 64), 1, 128, 3)
    run(3, (64, 64), 1, 77, 300)
    run(2, (32, 32), 1, 1000, 50)
    run(5, (16, 48, 32), 2, 333, 20)
    run(10, (64, 64, 64, 64), 4, 500, 20)
    run(10, (128, 128), 1, 1000, 20)
    # timing on the config-5 shape
    d, hls, N, K = 3, (64, 64), 10000, 4736
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
        print('QB_NO_TC=' + notc, 'logpost ms', med, 'evals/s', K / med * 1e3, 'TFLOP/s', K * 2.0 * N * desc.macs_per_point() / med / 1e9, flush=True)

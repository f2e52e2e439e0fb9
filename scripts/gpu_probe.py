"""Quick GPU probe: FMA peaks + first timings of kernels 1/2 on the benchmark shapes (development aid)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from quinn_b200 import ops                      # noqa: E402
from quinn_b200.netdesc import NetDesc, Layer   # noqa: E402


def mlp_desc(d, o, hls):
    widths = [d] + list(hls) + [o]
    layers, off = [], 0
    for l in range(len(widths) - 1):
        w = off
        off += widths[l] * widths[l + 1]
        b = off
        off += widths[l + 1]
        layers.append(Layer(widths[l], widths[l + 1], w, b, 'tanh' if l < len(widths) - 2 else 'identity', 0.0))
    return NetDesc(d, o, off, layers)


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(min(ts))


def main():
    res = {}
    print(torch.cuda.get_device_name(0), flush=True)
    for dt, name in ((torch.float32, 'f32'), (torch.float64, 'f64')):
        for v in (0, 1):
            res[f'fma_peak_{name}_v{v}_tflops'] = ops.fma_peak(dt, v, iters=20000) / 1e12
    print(json.dumps(res), flush=True)
    rs = np.random.RandomState(0)
    cases = [('c5', 3, (64, 64), 10000, 8192), ('c2', 2, (32, 32), 1000, 4096), ('c3', 10, (128, 128), 10000, 1024)]
    for name, d, hls, N, K in cases:
        desc = mlp_desc(d, 1, hls)
        x = rs.rand(N, d).astype(np.float32) * 2 - 1
        y = np.sin(x.sum(1, keepdims=True)).astype(np.float32)
        th = (0.2 * rs.randn(K, desc.n_params)).astype(np.float32)
        for dt, dn in ((torch.float32, 'f32'), (torch.float64, 'f64')):
            Kc = K if dt == torch.float32 else K // 8
            prob = ops.Problem(desc, x, y, 0.05, dtype=dt)
            tht = prob.theta(th[:Kc])
            S = desc.macs_per_point()
            fv = 2.0 * N * S
            fvg = 6.0 * N * S - 2.0 * N * desc.layers[0].n_in * desc.layers[0].n_out
            lp = torch.empty(Kc, dtype=torch.float64, device='cuda')
            med, best = timeit(lambda: ops.logpost(prob, tht, out=lp))
            r = dict(plan=prob.plan_info(Kc, False), ms=med, ms_best=best, evals_per_s=Kc / med * 1e3, tflops=Kc * fv / med / 1e9)
            res[f'{name}_{dn}_value'] = r
            print(name, dn, 'value', json.dumps(r), flush=True)
            g = torch.empty_like(tht)
            med, best = timeit(lambda: ops.logpost_grad(prob, tht, lp, g))
            r = dict(plan=prob.plan_info(Kc, True), ms=med, ms_best=best, evals_per_s=Kc / med * 1e3, tflops=Kc * fvg / med / 1e9)
            res[f'{name}_{dn}_grad'] = r
            print(name, dn, 'grad', json.dumps(r), flush=True)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'probe.json'), 'w') as f:
        json.dump(res, f, indent=1)


if __name__ == '__main__':
    main()

"""Time the pieces of the host-in / host-out AMCMC run at config-5 size (development aid)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from quinn_b200 import ops
from quinn_b200.mcmc import AMCMC, DeviceLogPost
spec = bench.workload_spec('c5')
desc = bench.mlp_desc(spec['d'], spec['hls'])
x, y = bench.make_data(spec)
K, P, steps = spec['K'], desc.n_params, 3
th = torch.from_numpy(bench.theta_init(spec, P, 0, K)).pin_memory()
prob = ops.Problem(desc, x, y, spec['sigma'], dtype=torch.float32)
sam = AMCMC(gamma=0.01, adapt='diag'); sam.setLogPost(DeviceLogPost(prob), None)
def T(fn, n=2):
    out = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); out.append(time.perf_counter() - t0); del r
    return out
print('pin 5.4GB', T(lambda: (torch.empty((K, 2, P), dtype=torch.float32, pin_memory=True), torch.empty((K, P), dtype=torch.float32, pin_memory=True)), 3))
d = torch.empty((K, P), dtype=torch.float32, device='cuda')
print('h2d 1.8GB', T(lambda: d.copy_(th, non_blocking=True)))
h = torch.empty((K, P), dtype=torch.float32).pin_memory()
print('d2h 1.8GB', T(lambda: h.copy_(d, non_blocking=True)))
for nsh in (4, 2, 8):
    sam._auto_shards = lambda *a, n=nsh: n
    print('run shards', nsh, T(lambda: sam.run(steps, th, seed=5, store_every=steps, verbose=False) if nsh > 1 else sam._finish(sam._run_fused(steps, th, 5, steps, None, 0, False), False, False), 3))

"""Phase timeline of the hot-shape AMCMC chain kernel (development aid; -DQB3_TRACE build, see tc3_trace.py).
   QB_LIB=quinn_b200/lib/libquinn_b200_trace.so python scripts/tc3_trace_chain.py [K] [steps]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
from gpu_probe import mlp_desc                  # noqa: E402
from quinn_b200 import ops, _lib                # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 296
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
N = 10000
rs = np.random.RandomState(0)
desc = mlp_desc(3, 1, (64, 64))
x = rs.rand(N, 3) * 2 * np.pi - np.pi
y = np.sin(x).sum(1, keepdims=True) + 0.05 * rs.randn(N, 1)
prob = ops.Problem(desc, x, y, 0.05, dtype=torch.float32)
th = prob.theta(rs.rand(K, desc.n_params))
st = ops.ChainState(prob, th)
am = ops.AmcmcState(st, gamma=0.01, t0=100, tadapt=1000, adapt='diag')
ops.amcmc_run(st, am, 3, None, seed=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.amcmc_run(st, am, steps, None, seed=1)
e1.record()
torch.cuda.synchronize()
print(f'K={K} steps={steps}: launch {e0.elapsed_time(e1) * 1e3:.1f} us -> {e0.elapsed_time(e1) * 1e3 / steps:.1f} us per step')
lib = _lib.load()
NB, NT, NE = 296, 82, 8
buf = np.zeros(NB * 17 * NT * NE, dtype=np.uint32)
sm = np.zeros(NB, dtype=np.uint32)
assert lib.qb_tc3_trace_dump(buf.ctypes.data_as(C.c_void_p), sm.ctypes.data_as(C.c_void_p)) == 0
buf = buf.reshape(NB, 17, NT, NE).astype(np.int64)
nb = min(K, NB)
M = 0xFFFFFFFF
ph = buf[:nb, 0, 81]                  # warp 0 of every block: stamps of the LAST step
k0 = buf[:nb, 0, 80]                  # kernel-level stamps
e = buf[:nb, 0]
f = lambda v: f'{(v & M).mean():9.0f}'
print('kernel start: tmem alloc + init', f(k0[:, 1] - k0[:, 0]), ' state load', f(k0[:, 2] - k0[:, 1]), ' block lifetime', f(k0[:, 3] - k0[:, 0]),
      ' per step', f((k0[:, 3] - k0[:, 0]) // steps))
print('last step (cycles, warp 0, mean over blocks):')
print('  per-element phase          ', f(ph[:, 1] - ph[:, 0]))
print('    start -> loads issued    ', f(ph[:, 5] - ph[:, 0]), ' group 0', f(ph[:, 6] - ph[:, 5]), ' group 1', f(ph[:, 7] - ph[:, 6]), ' groups 2..4', f(ph[:, 1] - ph[:, 7]))
print('  barrier + staging          ', f(ph[:, 2] - ph[:, 1]))
print('  eval call                  ', f(ph[:, 3] - ph[:, 2]))
print('    staging end -> EPI0(0)   ', f(e[:, 0, 0] - ph[:, 2]))
print('    EPI0(0) start -> arrive  ', f(e[:, 0, 4] - e[:, 0, 0]))
print('    EPI0(0) -> EPI0(10) start', f(e[:, 10, 0] - e[:, 0, 0]))
print('    EPI0(10) -> EPI0(70)     ', f(e[:, 70, 0] - e[:, 10, 0]), ' = per tile', f((e[:, 70, 0] - e[:, 10, 0]) // 60))
print('    EPI0(70) -> EPI1(78) end ', f(e[:, 78, 7] - e[:, 70, 0]))
print('    EPI1(78) end -> returned ', f(ph[:, 3] - e[:, 78, 7]))
print('  accept + bookkeeping       ', f(ph[:, 4] - ph[:, 3]))
print('  after last step -> stored  ', f(k0[:, 3] - ph[:, 4]))

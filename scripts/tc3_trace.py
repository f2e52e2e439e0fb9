"""Phase timeline of the warp-specialised value kernel (development aid; needs the -DQB3_TRACE build of
qb_value_tc3.cu linked as quinn_b200/lib/libquinn_b200_trace.so).  One wave of kernel 1 at the config-5 shape with 2
blocks per SM; prints, for the blocks of one SM, the per-tile phase durations of every warp (SM clock cycles).
   QB_LIB=quinn_b200/lib/libquinn_b200_trace.so python scripts/tc3_trace.py [K]"""
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

os.environ['QB_SPLIT'] = '1'
K = int(sys.argv[1]) if len(sys.argv) > 1 else 296
N, T = 10112, 79
rs = np.random.RandomState(0)
desc = mlp_desc(3, 1, (64, 64))
x = rs.rand(N, 3) * 2 - 1
y = np.sin(x.sum(1, keepdims=True))
prob = ops.Problem(desc, x, y, 0.05, dtype=torch.float32)
th = prob.theta(0.2 * rs.randn(K, desc.n_params))
lp = torch.empty(K, dtype=torch.float64, device='cuda')
for _ in range(3):
    ops.logpost(prob, th, lp)
torch.cuda.synchronize()
lib = _lib.load()
NB, NT, NE = 296, 82, 8
buf = np.zeros(NB * 17 * NT * NE, dtype=np.uint32)
sm = np.zeros(NB, dtype=np.uint32)
rc = lib.qb_tc3_trace_dump(buf.ctypes.data_as(C.c_void_p), sm.ctypes.data_as(C.c_void_p))
assert rc == 0
buf = buf.reshape(NB, 17, NT, NE).astype(np.int64)
sm = sm[:min(K, NB)]
sm0 = sm[0]
blocks = [b for b in range(min(K, NB)) if sm[b] == sm0]
print('blocks on SM', sm0, ':', blocks)
names = ['e0.start', 'e0.ld', 'e0.comp', 'e0.d1f', 'e0.arr', 'e1.start', 'e1.ld', 'e1.end']
t0 = min(buf[b, w, 0, 0] for b in blocks for w in range(8))
for b in blocks:
    print(f'--- block {b}: per-warp mean phase durations over tiles 10..70 (cycles)')
    for w in range(8):
        e = buf[b, w]          # [tile, ev]
        u = np.arange(10, 70)
        d_wait0 = (e[u, 1] - e[u, 0]).mean()       # d0_full wait + tmem load
        d_comp0 = (e[u, 2] - e[u, 1]).mean()       # sigmoid + split
        d_wait1 = (e[u, 3] - e[u, 2]).mean()       # d1_full wait
        d_st = (e[u, 4] - e[u, 3]).mean()          # tmem store + arrive
        d_gap = (e[u - 1, 5] - e[u, 4]).mean()     # -> EPI1(u-1) start
        d_ld1 = (e[u - 1, 6] - e[u - 1, 5]).mean()
        d_comp1 = (e[u - 1, 7] - e[u - 1, 6]).mean()
        per = (e[u + 1, 0] - e[u, 0]).mean()
        print(f'  warp {w}: period {per:7.0f} | EPI0 wait+ld {d_wait0:6.0f} comp {d_comp0:6.0f} d1f-wait {d_wait1:6.0f} st+arr {d_st:5.0f} | EPI1 ld {d_ld1:5.0f} comp {d_comp1:6.0f}')
    e = buf[b, 8]
    u = np.arange(10, 70)
    print(f'  issuer: period {(e[u + 1, 0] - e[u, 0]).mean():7.0f} | mma0 issue {(e[u, 1] - e[u, 0]).mean():5.0f} d1free wait {(e[u, 2] - e[u, 1]).mean():5.0f} mma1 issue {(e[u, 3] - e[u, 2]).mean():5.0f} rest {(e[u + 1, 0] - e[u, 3]).mean():6.0f}')
    # skew: arrival times on a_ready(u) relative to the issuer's wake-up
    arr = np.stack([buf[b, w, 10:70, 4] for w in range(8)])       # [warp, tile]
    wake = buf[b, 8, 10:70, 0]
    print(f'  a_ready: last arrival -> issuer wake {np.mean(wake - arr.max(0)):6.0f}; first -> last arrival {np.mean(arr.max(0) - arr.min(0)):6.0f}')
    d1f_need = np.stack([buf[b, w, 11:71, 2] for w in range(8)])  # when EPI0(u+1) wants d1_full(u)
    mma1_done_proxy = buf[b, 8, 10:70, 3]
    print(f'  MMA1(u) issued (issuer stamp) -> EPI0(u+1) stores unblock: mean {np.mean(np.stack([buf[b, w, 11:71, 3] for w in range(8)]).min(0) - mma1_done_proxy):6.0f}; warps reach the d1f wait {np.mean(d1f_need.min(0) - mma1_done_proxy):6.0f} (first) {np.mean(d1f_need.max(0) - mma1_done_proxy):6.0f} (last) after it')
print('total eval cycles (block 0, warp 0):', buf[blocks[0], 0, T - 1, 7] - buf[blocks[0], 0, 0, 0])

"""Phase timeline of the 128-wide tensor-core gradient kernel (development aid; needs the -DQB_TG8_TRACE build of
qb_grad_tc128.cu linked as quinn_b200/lib/libquinn_b200_trace.so: scripts/build_trace_lib.sh).  One wave of kernel 2 at the
config-4 shape; prints the per-tile phase durations of every warp of block 0 (SM clock cycles).
   QB_LIB=quinn_b200/lib/libquinn_b200_trace.so python scripts/tg8_trace.py [H [K]]"""
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
H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
K = int(sys.argv[2]) if len(sys.argv) > 2 else 148
N = 128 * 44
d_in = 10 if H == 128 else 3
rs = np.random.RandomState(0)
desc = mlp_desc(d_in, 1, (H, H))
x = rs.rand(N, d_in) * 2 - 1
y = np.sin(x.sum(1, keepdims=True))
prob = ops.Problem(desc, x, y, 0.05, dtype=torch.float32)
th = prob.theta(0.2 * rs.randn(K, desc.n_params))
lp = torch.empty(K, dtype=torch.float64, device='cuda')
g = torch.empty_like(th)
for _ in range(3):
    ops.logpost_grad(prob, th, lp, g)
torch.cuda.synchronize()
lib = _lib.load()
NB, NW, NT, NE = 8, 17, 48, 12
buf = np.zeros(NB * NW * NT * NE, dtype=np.uint32)
lib.qb_tg8_trace_dump.restype = C.c_int
assert lib.qb_tg8_trace_dump(buf.ctypes.data_as(C.c_void_p)) == 0
buf = buf.reshape(NB, NW, NT, NE).astype(np.int64)
u = np.arange(8, 40)
for b in (0, 1):
    print(f'--- block {b}: mean phase durations over tiles 8..39 (cycles)')
    for w in range(H // 8):
        e = buf[b, w]
        per = (e[u + 1, 0] - e[u, 0]).mean()
        d = [(e[u, i + 1] - e[u, i]).mean() for i in range(9)]
        print(f'  warp {w:2d}: period {per:7.0f} | wait f {d[0]:6.0f} | EPI1a {d[1]:5.0f} sync+zf {d[2]:5.0f} EPI1b {d[3]:5.0f} | pub, wait l, EPIL {d[4]:6.0f} | '
              f'(wait l {(e[u, 10] - e[u, 4]).mean():5.0f}) wait b {d[5]:5.0f} EPI0 {d[6]:5.0f} | wait w {d[7]:6.0f} | unpark + pub {d[8]:5.0f}')
    e = buf[b, H // 8]
    print(f'  issuer : period {(e[u + 1, 0] - e[u, 0]).mean():7.0f} | issue BWD+DW1 {(e[u, 1] - e[u, 0]).mean():5.0f} wait rdy(A) {(e[u, 2] - e[u, 1]).mean():6.0f} '
          f'issue FWD+DW0 {(e[u, 3] - e[u, 2]).mean():5.0f} wait rdy(B) {(e[u + 1, 0] - e[u, 3]).mean():6.0f}')
    # tensor-pipe view: from the issuer's stamps to the first warp that sees each barrier
    w0 = buf[b, 0]
    print(f'  FWD+DW0 issue start -> warp0 sees f(t+1): {(w0[u + 1, 1] - e[u, 2]).mean():6.0f};  BWD issue start -> warp0 sees b: {(w0[u, 6] - e[u, 0]).mean():6.0f};'
          f'  -> sees w: {(w0[u, 8] - e[u, 0]).mean():6.0f}')

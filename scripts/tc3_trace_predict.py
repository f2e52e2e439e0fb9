"""Phase timeline of the 128-wide warp-specialised predict kernel (development aid; -DQB3_TRACE build).
   QB_LIB=quinn_b200/lib/libquinn_b200_trace.so python scripts/tc3_trace_predict.py"""
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

d, H, M, N = 10, 128, 148, 128 * 79 * 16
rs = np.random.RandomState(0)
desc = mlp_desc(d, 1, (H, H))
xx = torch.as_tensor(rs.rand(N, d), dtype=torch.float32, device='cuda')
th = torch.as_tensor((2 * rs.rand(M, desc.n_params) - 1) / np.sqrt(H), dtype=torch.float32, device='cuda')
for _ in range(3):
    ops.predict(desc, th, xx, dtype=torch.float32, want_out=True, want_moments=False)
torch.cuda.synchronize()
lib = _lib.load()
NB, NW, NT, NE = 296, 17, 82, 8
buf = np.zeros(NB * NW * NT * NE, dtype=np.uint32)
sm = np.zeros(NB, dtype=np.uint32)
assert lib.qb_tc3_trace_dump(buf.ctypes.data_as(C.c_void_p), sm.ctypes.data_as(C.c_void_p)) == 0
buf = buf.reshape(NB, NW, NT, NE).astype(np.int64)
b = 0
T = int((buf[b, 0, :, 0] != 0).sum())
print('tiles traced in block 0:', T)
u = np.arange(5, T - 5)
for w in range(16):
    e = buf[b, w]
    print(f'  warp {w:2d}: period {(e[u + 1, 0] - e[u, 0]).mean():7.0f} | EPI0 wait+ld {(e[u, 1] - e[u, 0]).mean():6.0f} comp {(e[u, 2] - e[u, 1]).mean():6.0f} d1f-wait {(e[u, 3] - e[u, 2]).mean():6.0f} st+arr {(e[u, 4] - e[u, 3]).mean():5.0f} | EPI1 ld {(e[u - 1, 6] - e[u - 1, 5]).mean():5.0f} comp {(e[u - 1, 7] - e[u - 1, 6]).mean():6.0f}')
e = buf[b, 16]
print(f'  issuer: period {(e[u + 1, 0] - e[u, 0]).mean():7.0f} | mma0 issue {(e[u, 1] - e[u, 0]).mean():5.0f} d1free wait {(e[u, 2] - e[u, 1]).mean():5.0f} mma1 issue {(e[u, 3] - e[u, 2]).mean():5.0f} rest (x staging + wait) {(e[u + 1, 0] - e[u, 3]).mean():6.0f}')
arr = np.stack([buf[b, w, u, 4] for w in range(16)])
print(f'  a_ready: last arrival -> issuer wake {np.mean(e[u, 0] - arr.max(0)):6.0f}; first -> last arrival {np.mean(arr.max(0) - arr.min(0)):6.0f}')

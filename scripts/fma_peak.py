import sys; sys.path.insert(0,'.')
import torch
from quinn_b200 import ops
for v in (0,1,2):
    print('f32 variant',v, ops.fma_peak(torch.float32, v, iters=20000)/1e12, 'TFLOP/s')

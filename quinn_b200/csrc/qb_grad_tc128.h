// Host interface of the 128-wide tensor-core gradient translation unit (qb_grad_tc128.cu); called from qb_kernels.cu.
#pragma once
#include <cuda_runtime.h>
#include "quinn_b200.h"
#include "qb_tg8_plan.h"

template <typename T> struct EvalArgs;

// eligibility + plan of the 128-wide tcgen05 gradient path (fp32, in <= 15 -> 128 -> 128 -> 1, tanh); false: not eligible
bool qb_tg8_make_plan(const qb_net_t* net, int dtype, QbTg8Plan* tp);
// scratch: 8 bytes of device memory (max |x|, max |y| of the launch), written by a small kernel ahead of the evaluation
cudaError_t qb_tg8_launch_eval(const QbTg8Plan& tp, const EvalArgs<float>& a, void* scratch, dim3 grid, cudaStream_t st);
enum { QB_TG8_SCRATCH_BYTES = 256 };

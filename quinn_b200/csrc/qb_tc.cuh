// Tensor-core evaluation of the data term for fp32 MLPs: tcgen05.mma kind::tf32 with the operands split in two
// ("3xTF32": a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo, products accumulated in fp32 in tensor memory), which keeps
// fp32-level accuracy (relative error of a product ~2^-21) while the 64x64 hidden-layer GEMMs leave the CUDA cores.
//
// One thread block = one warpgroup of 128 threads = one tile of 128 data points; thread t owns point t:
//   layer 0 (n_in <= 15):  CUDA cores, thread-per-point; tanh; the result is split into hi/lo and written with
//                          tcgen05.st to tensor memory as the A operand (lane = point, column = unit);
//   hidden layers 1..L-2:  D[128 x n_out] = A[128 x n_in] * W^T, A from tensor memory, W (hi and lo) staged once per
//                          parameter vector in shared memory in the K-major no-swizzle canonical layout (W is stored
//                          (n_out, n_in) row-major in theta = already K-major).  One elected thread issues
//                          3 * n_in/8 MMAs and a tcgen05.commit to an mbarrier; everybody waits on the mbarrier, reads
//                          D back with tcgen05.ld, adds the bias, applies the activation and either writes the next
//                          A operand or, for the last hidden layer, feeds
//   last layer (n_out<=4): a per-thread dot product with weights broadcast from shared memory, then the residual.
// Tensor-memory columns: [0,Kmax) A_hi, [Kmax,2Kmax) A_lo, [2Kmax, 2Kmax+Nmax) D.
// The tensor pipe of one block overlaps with the CUDA-core phases of the other block(s) resident on the SM.
#pragma once
#include <stdint.h>
#include "qb_plan.h"

struct QbTcLayer {
    int n_in, n_out, w_off, b_off, act;
    int bhi, blo;        // byte offsets in dynamic shared memory of the split weight tiles (canonical layout)
    int bias;            // float index (in the float area) of the n_out biases
};
struct QbTcPlan {
    int n_layers;                       // layers 1 .. n_layers-2 run on the tensor cores
    int in_dim, ni, out_dim, n_params;  // ni: padded input width (4, 8 or 16; slot in_dim carries the bias)
    int h0, kl;                         // width after layer 0; n_in of the last layer
    int act0, act_last, final_exp;
    int w0, wl, bl;                     // float indices: layer-0 rows [h0][ni], last-layer W [out][kl], last bias
    int w0_off, b0_off, wl_off, bl_off; // offsets in theta (b*_off < 0: no bias)
    int fl_base;                        // byte offset of the float area
    int a_lo_col, d_col, tmem_cols;
    int smem_bytes;
    QbTcLayer L[QB_MAX_LAYERS];
};

#ifdef __CUDACC__
enum { QB_TC_RED_BYTES = 320, QB_TC_BAR_OFF = 320, QB_TC_SLOT_OFF = 328, QB_TC_HDR_BYTES = 384 };

struct QbTcCtx { uint32_t tmem, bar, phase; };

__device__ __forceinline__ uint32_t qb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define QB_R16(v) "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), \
                  "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
#define QB_W16(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
#define QB_RW16(v) "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), \
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])

__device__ __forceinline__ void qb_tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), QB_R16(v) : "memory");
}
__device__ __forceinline__ void qb_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : QB_W16(v) : "r"(taddr) : "memory");
}
// the registers are operands of the wait so that no use of them can be scheduled before it
__device__ __forceinline__ void qb_tmem_ld_wait16(uint32_t (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : QB_RW16(v) :: "memory");
}
__device__ __forceinline__ void qb_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void qb_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void qb_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// bounded wait: a lost commit traps instead of hanging the GPU
__device__ __forceinline__ void qb_mbar_wait(uint32_t bar, uint32_t parity) {
    for (int it = 0; it < (1 << 24); ++it) {
        uint32_t ok;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}

// all threads; allocates tensor memory and initialises the mbarrier
__device__ __forceinline__ void qb_tc_init(const QbTcPlan& tp, unsigned char* smem, QbTcCtx& cx) {
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(qb_smem_u32(smem + QB_TC_SLOT_OFF)), "r"((uint32_t)tp.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(qb_smem_u32(smem + QB_TC_BAR_OFF)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    qb_tc_fence_before();
    __syncthreads();
    qb_tc_fence_after();
    cx.tmem = *reinterpret_cast<volatile uint32_t*>(smem + QB_TC_SLOT_OFF);
    cx.bar = qb_smem_u32(smem + QB_TC_BAR_OFF);
    cx.phase = 0;
}
__device__ __forceinline__ void qb_tc_fini(const QbTcPlan& tp, const QbTcCtx& cx) {
    qb_tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(cx.tmem), "r"((uint32_t)tp.tmem_cols) : "memory");
}

__device__ __forceinline__ float qb_tf32_hi(float w) { return __uint_as_float(__float_as_uint(w) & 0xFFFFE000u); }

// flat theta (global) -> shared: layer-0 rows, split hidden weights in the canonical K-major layout, biases, last layer.
// Canonical layout (no swizzle): core matrix = 8 rows (n) x 16 bytes (4 k), stored as 128 contiguous bytes; core
// matrices adjacent along k (LBO = 128 B), groups of 8 rows SBO = 128*(K/4) B apart => float index
// e = ((n/8)*(K/4) + k/4)*32 + (n%8)*4 + k%4.
__device__ __forceinline__ void qb_tc_stage(const QbTcPlan& tp, unsigned char* smem, const float* __restrict__ theta) {
    float* F = reinterpret_cast<float*>(smem + tp.fl_base);
    const int tid = threadIdx.x, nt = blockDim.x;
    const float fold = 2.8853900817779268f;
    {
        const float s0 = tp.act0 == QB_ACT_TANH ? fold : 1.0f;
        for (int e = tid; e < tp.h0 * tp.ni; e += nt) {
            const int j = e / tp.ni, q = e - j * tp.ni;
            float v = 0.0f;
            if (q < tp.in_dim) v = theta[tp.w0_off + j * tp.in_dim + q] * s0;
            else if (q == tp.in_dim && tp.b0_off >= 0) v = theta[tp.b0_off + j] * s0;
            F[tp.w0 + e] = v;
        }
    }
    for (int l = 1; l < tp.n_layers - 1; ++l) {
        const QbTcLayer& L = tp.L[l];
        const int K = L.n_in, N = L.n_out, kc4 = K >> 2;
        const float sc = L.act == QB_ACT_TANH ? fold : 1.0f;
        float* hi = reinterpret_cast<float*>(smem + L.bhi);
        float* lo = reinterpret_cast<float*>(smem + L.blo);
        for (int e = tid; e < N * K; e += nt) {
            const int q = e & 3, r = (e >> 2) & 7, g = e >> 5;
            const int n8 = g / kc4, kc = g - n8 * kc4;
            const float w = theta[L.w_off + (n8 * 8 + r) * K + kc * 4 + q] * sc;
            const float h = qb_tf32_hi(w);
            hi[e] = h;
            lo[e] = w - h;
        }
        for (int j = tid; j < N; j += nt) F[L.bias + j] = L.b_off >= 0 ? theta[L.b_off + j] * sc : 0.0f;
    }
    {
        const float sl = tp.act_last == QB_ACT_TANH ? fold : 1.0f;
        for (int e = tid; e < tp.out_dim * tp.kl; e += nt) F[tp.wl + e] = theta[tp.wl_off + e] * sl;
        for (int j = tid; j < tp.out_dim; j += nt) F[tp.bl + j] = tp.bl_off >= 0 ? theta[tp.bl_off + j] * sl : 0.0f;
    }
    // the tensor core reads shared memory through the async proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// one elected thread: D = A_lo*Bhi^T + A_hi*Blo^T + A_hi*Bhi^T, then commit to the mbarrier
__device__ __forceinline__ void qb_tc_issue(const QbTcPlan& tp, const QbTcLayer& L, const QbTcCtx& cx, unsigned char* smem) {
    const uint32_t K = L.n_in, N = L.n_out;
    // instruction descriptor: D fp32, A and B tf32, both K-major, N>>3 at bit 17, M=128 (>>4) at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sbo = 128u * (K >> 2);
    const uint64_t dfix = ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
    const uint32_t bhi = qb_smem_u32(smem + L.bhi), blo = qb_smem_u32(smem + L.blo);
    const uint32_t d = cx.tmem + tp.d_col;
    uint32_t acc = 0;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a = cx.tmem + (pass == 0 ? tp.a_lo_col : 0);
        const uint32_t b = pass == 1 ? blo : bhi;
#pragma unroll 4
        for (uint32_t s = 0; s < K / 8; ++s) {
            const uint64_t desc = dfix | (uint64_t)(((b + s * 256u) >> 4) & 0x3FFFu);
            asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p; }"
                         :: "r"(d), "r"(a + s * 8u), "l"(desc), "r"(idesc), "r"(acc) : "memory");
            acc = 1;
        }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(cx.bar) : "memory");
}

template <int ACT> __device__ __forceinline__ float qb_tc_act(float z) {
    if (ACT == QB_ACT_TANH) return qb_tanh_prescaled(z);
    if (ACT == QB_ACT_RELU) return fmaxf(z, 0.0f);
    return z;
}

__device__ __forceinline__ void qb_tc_split_store(uint32_t t_hi, uint32_t t_lo, const float (&h)[16]) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float hh = qb_tf32_hi(h[i]);
        hi[i] = __float_as_uint(hh);
        lo[i] = __float_as_uint(h[i] - hh);
    }
    qb_tmem_st16(t_hi, hi);
    qb_tmem_st16(t_lo, lo);
}

// layer 0 on the CUDA cores: h = act(W0 x + b0) for this thread's point, written as the first A operand
template <int NI, int ACT>
__device__ __forceinline__ void qb_tc_layer0(const QbTcPlan& tp, const float* F, uint32_t tl, const float (&xr)[NI]) {
    const float4* W = reinterpret_cast<const float4*>(F + tp.w0);
#pragma unroll 1
    for (int c = 0; c < tp.h0; c += 16) {
        float h[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float z = 0.0f;
#pragma unroll
            for (int q4 = 0; q4 < NI / 4; ++q4) {
                const float4 w = W[(c + i) * (NI / 4) + q4];
                z = fmaf(w.x, xr[q4 * 4 + 0], z); z = fmaf(w.y, xr[q4 * 4 + 1], z);
                z = fmaf(w.z, xr[q4 * 4 + 2], z); z = fmaf(w.w, xr[q4 * 4 + 3], z);
            }
            h[i] = qb_tc_act<ACT>(z);
        }
        qb_tc_split_store(tl + c, tl + tp.a_lo_col + c, h);
    }
}

// hidden-layer epilogue: D -> +bias -> act -> next A operand
template <int ACT>
__device__ __forceinline__ void qb_tc_epi_mid(const QbTcPlan& tp, const QbTcLayer& L, const float* F, uint32_t tl) {
    const float4* B4 = reinterpret_cast<const float4*>(F + L.bias);
#pragma unroll 1
    for (int c = 0; c < L.n_out; c += 16) {
        uint32_t v[16];
        qb_tmem_ld16(tl + tp.d_col + c, v);
        qb_tmem_ld_wait16(v);
        float h[16];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
            const float4 b = B4[(c >> 2) + i4];
            h[i4 * 4 + 0] = qb_tc_act<ACT>(__uint_as_float(v[i4 * 4 + 0]) + b.x);
            h[i4 * 4 + 1] = qb_tc_act<ACT>(__uint_as_float(v[i4 * 4 + 1]) + b.y);
            h[i4 * 4 + 2] = qb_tc_act<ACT>(__uint_as_float(v[i4 * 4 + 2]) + b.z);
            h[i4 * 4 + 3] = qb_tc_act<ACT>(__uint_as_float(v[i4 * 4 + 3]) + b.w);
        }
        qb_tc_split_store(tl + c, tl + tp.a_lo_col + c, h);
    }
}

// last hidden layer's epilogue fused with the narrow output layer: yacc[o] += W_last[o][j] * act(D[j] + b[j])
template <int ACT>
__device__ __forceinline__ void qb_tc_epi_last(const QbTcPlan& tp, const QbTcLayer& L, const float* F, uint32_t tl,
                                               float (&yacc)[4]) {
    const float4* B4 = reinterpret_cast<const float4*>(F + L.bias);
    const float* WL = F + tp.wl;
    const int od = tp.out_dim, kl = tp.kl;
#pragma unroll 1
    for (int c = 0; c < L.n_out; c += 16) {
        uint32_t v[16];
        qb_tmem_ld16(tl + tp.d_col + c, v);
        qb_tmem_ld_wait16(v);
        float h[16];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
            const float4 b = B4[(c >> 2) + i4];
            h[i4 * 4 + 0] = qb_tc_act<ACT>(__uint_as_float(v[i4 * 4 + 0]) + b.x);
            h[i4 * 4 + 1] = qb_tc_act<ACT>(__uint_as_float(v[i4 * 4 + 1]) + b.y);
            h[i4 * 4 + 2] = qb_tc_act<ACT>(__uint_as_float(v[i4 * 4 + 2]) + b.z);
            h[i4 * 4 + 3] = qb_tc_act<ACT>(__uint_as_float(v[i4 * 4 + 3]) + b.w);
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            if (o < od) {
                const float4* w4 = reinterpret_cast<const float4*>(WL + o * kl + c);
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    const float4 w = w4[i4];
                    yacc[o] = fmaf(w.x, h[i4 * 4 + 0], yacc[o]); yacc[o] = fmaf(w.y, h[i4 * 4 + 1], yacc[o]);
                    yacc[o] = fmaf(w.z, h[i4 * 4 + 2], yacc[o]); yacc[o] = fmaf(w.w, h[i4 * 4 + 3], yacc[o]);
                }
            }
        }
    }
}

// forward pass of the tile of 128 points starting at p0 (thread t = point p0+t); leaves the network output of this
// thread's point in yout[0..out_dim)
template <int NI>
__device__ __forceinline__ void qb_tc_forward_tile(const QbTcPlan& tp, QbTcCtx& cx, unsigned char* smem,
                                                   const float* __restrict__ x, int64_t p, bool live, float (&yout)[4]) {
    const float* F = reinterpret_cast<const float*>(smem + tp.fl_base);
    const uint32_t tl = cx.tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
    float xr[NI];
#pragma unroll
    for (int q = 0; q < NI; ++q) {
        xr[q] = 0.0f;
        if (q < tp.in_dim) { if (live) xr[q] = __ldg(x + p * tp.in_dim + q); }
        else if (q == tp.in_dim) xr[q] = 1.0f;                     // bias slot
    }
    switch (tp.act0) {
        case QB_ACT_TANH: qb_tc_layer0<NI, QB_ACT_TANH>(tp, F, tl, xr); break;
        case QB_ACT_RELU: qb_tc_layer0<NI, QB_ACT_RELU>(tp, F, tl, xr); break;
        default: qb_tc_layer0<NI, QB_ACT_IDENTITY>(tp, F, tl, xr); break;
    }
    float yacc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const int last_tc = tp.n_layers - 2;
    for (int l = 1; l <= last_tc; ++l) {
        const QbTcLayer& L = tp.L[l];
        // A (and the reads of D by the previous tile / layer) are complete in every thread before the MMAs start
        qb_tmem_st_wait();
        qb_tc_fence_before();
        __syncthreads();
        if (threadIdx.x == 0) {
            qb_tc_fence_after();
            qb_tc_issue(tp, L, cx, smem);
        }
        qb_mbar_wait(cx.bar, cx.phase);
        cx.phase ^= 1u;
        qb_tc_fence_after();
        if (l < last_tc) {
            switch (L.act) {
                case QB_ACT_TANH: qb_tc_epi_mid<QB_ACT_TANH>(tp, L, F, tl); break;
                case QB_ACT_RELU: qb_tc_epi_mid<QB_ACT_RELU>(tp, L, F, tl); break;
                default: qb_tc_epi_mid<QB_ACT_IDENTITY>(tp, L, F, tl); break;
            }
        } else {
            switch (L.act) {
                case QB_ACT_TANH: qb_tc_epi_last<QB_ACT_TANH>(tp, L, F, tl, yacc); break;
                case QB_ACT_RELU: qb_tc_epi_last<QB_ACT_RELU>(tp, L, F, tl, yacc); break;
                default: qb_tc_epi_last<QB_ACT_IDENTITY>(tp, L, F, tl, yacc); break;
            }
        }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        float v = 0.0f;
        if (o < tp.out_dim) {
            v = yacc[o] + F[tp.bl + o];
            if (tp.act_last == QB_ACT_TANH) v = qb_tanh_prescaled(v);
            else if (tp.act_last == QB_ACT_RELU) v = fmaxf(v, 0.0f);
            if (tp.final_exp) v = expf(v);
        }
        yout[o] = v;
    }
}

// sum of squared residuals over points [n0, n1) for the staged parameter vector (block-wide result)
template <int NI>
__device__ __forceinline__ double qb_tc_eval_ni(const QbTcPlan& tp, QbTcCtx& cx, unsigned char* smem,
                                                const float* __restrict__ x, const float* __restrict__ y,
                                                int64_t n0, int64_t n1) {
    float ssq = 0.0f;
    for (int64_t p0 = n0; p0 < n1; p0 += 128) {
        const int64_t p = p0 + threadIdx.x;
        const bool live = p < n1;
        float yv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int o = 0; o < 4; ++o) if (o < tp.out_dim && live) yv[o] = __ldg(y + p * tp.out_dim + o);
        float yo[4];
        qb_tc_forward_tile<NI>(tp, cx, smem, x, p, live, yo);
        if (live) {
#pragma unroll
            for (int o = 0; o < 4; ++o) if (o < tp.out_dim) { const float r = yv[o] - yo[o]; ssq = fmaf(r, r, ssq); }
        }
    }
    return qb_block_sum((double)ssq, reinterpret_cast<double*>(smem));
}

__device__ __forceinline__ double qb_tc_eval(const QbTcPlan& tp, QbTcCtx& cx, unsigned char* smem,
                                             const float* __restrict__ x, const float* __restrict__ y,
                                             int64_t n0, int64_t n1) {
    if (tp.ni == 4) return qb_tc_eval_ni<4>(tp, cx, smem, x, y, n0, n1);
    if (tp.ni == 8) return qb_tc_eval_ni<8>(tp, cx, smem, x, y, n0, n1);
    return qb_tc_eval_ni<16>(tp, cx, smem, x, y, n0, n1);
}
#endif  // __CUDACC__

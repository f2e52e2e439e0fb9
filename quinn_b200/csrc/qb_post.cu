// quinn_b200: device-side chain / ensemble post-processing so that neither the (M, N*, o)
// predictive array nor the (steps, K) records have to visit the host: per-row moments (R-hat inputs), quantiles over
// the member axis (quinn/utils/stats.py:8-32 get_stats), effective sample size of a monitored scalar per chain, and the
// column mean of squares that turns per-point gradients into the diagonal Fisher (quinn/nns/nnwrap.py:204-229).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "quinn_b200.h"

// error text and launch counter live in qb_kernels.cu (qb_last_error / qb_launch_count)
void qb_internal_set_error(const char* msg);
void qb_internal_count_launches(int n);
static int pfail(const char* m) { qb_internal_set_error(m); return -1; }
#define QP_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { char b_[256]; snprintf(b_, sizeof(b_), "%s failed: %s", #call, cudaGetErrorString(e_)); qb_internal_set_error(b_); return -2; } } while (0)
#define QP_LAUNCHED() do { QP_CUDA(cudaGetLastError()); qb_internal_count_launches(1); } while (0)

__device__ __forceinline__ double qp_block_sum(double v, double* red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        const int nw = (blockDim.x + 31) >> 5;
        for (int w = 0; w < nw; ++w) s += red[w];
        red[32] = s;
    }
    __syncthreads();
    return red[32];
}

// ---- per-row mean and variance (ddof = 1) of x[K, n] (doubles): the per-chain moments R-hat needs
__global__ void __launch_bounds__(256) k_row_moments(const double* x, long long n, double* mean, double* var) {
    __shared__ double red[40];
    const double* r = x + (long long)blockIdx.x * n;
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) s += r[i];
    const double mu = qp_block_sum(s, red) / (double)n;
    double q = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) { const double d = r[i] - mu; q += d * d; }
    q = qp_block_sum(q, red);
    if (threadIdx.x == 0) { mean[blockIdx.x] = mu; var[blockIdx.x] = n > 1 ? q / (double)(n - 1) : NAN; }
}
extern "C" int qb_row_moments(const double* x, int64_t K, int64_t n, double* mean, double* var, void* stream) {
    if (!x || !mean || !var || K < 1 || n < 1) return pfail("bad argument to qb_row_moments");
    k_row_moments<<<(unsigned)K, 256, 0, (cudaStream_t)stream>>>(x, n, mean, var);
    QP_LAUNCHED();
    return 0;
}

// ---- effective sample size of x[K, n] per row: n / tau, tau = -1 + 2 * sum_m (rho_2m + rho_2m+1) over the initial
// positive sequence (Geyer 1992), autocorrelations with the biased 1/n normalisation
__global__ void __launch_bounds__(256) k_ess(const double* x, long long n, long long max_lag, double* ess, double* tau_out) {
    __shared__ double red[40];
    const double* r = x + (long long)blockIdx.x * n;
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) s += r[i];
    const double mu = qp_block_sum(s, red) / (double)n;
    double q = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) { const double d = r[i] - mu; q += d * d; }
    const double c0 = qp_block_sum(q, red);
    double tau = -1.0;
    if (c0 > 0.0) {
        double prev = 1.0;                 // rho_0
        for (long long t = 1; t <= max_lag; ++t) {
            double a = 0.0;
            for (long long i = threadIdx.x; i + t < n; i += blockDim.x) a += (r[i] - mu) * (r[i + t] - mu);
            const double rho = qp_block_sum(a, red) / c0;
            if (t & 1) {                   // pair (rho_{t-1}, rho_t) complete
                const double pair = prev + rho;
                if (!(pair > 0.0)) break;
                tau += 2.0 * pair;
            } else {
                prev = rho;
            }
        }
    } else {
        tau = NAN;
    }
    if (threadIdx.x == 0) {
        if (tau_out) tau_out[blockIdx.x] = tau;
        ess[blockIdx.x] = (tau > 0.0) ? fmin((double)n / tau, (double)n * 1e6) : NAN;
    }
}
extern "C" int qb_ess(const double* x, int64_t K, int64_t n, int64_t max_lag, double* ess, double* tau, void* stream) {
    if (!x || !ess || K < 1 || n < 2) return pfail("bad argument to qb_ess");
    if (max_lag <= 0 || max_lag > n - 1) max_lag = n - 1;
    if ((max_lag & 1) == 0) max_lag -= 1;                  // an odd last lag completes the last pair
    if (max_lag < 1) max_lag = 1;
    k_ess<<<(unsigned)K, 256, 0, (cudaStream_t)stream>>>(x, n, max_lag, ess, tau);
    QP_LAUNCHED();
    return 0;
}

// ---- column mean of squares of g[K, P]: out[p] = mean_k g[k,p]^2 (double), coalesced over p
template <typename T>
__global__ void k_colsq_mean(const T* g, long long K, long long P, double* out) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    double s = 0.0;
    for (long long k = 0; k < K; ++k) { const double v = (double)g[k * P + p]; s += v * v; }
    out[p] = s / (double)K;
}
extern "C" int qb_colsq_mean(int dtype, const void* g, int64_t K, int64_t P, double* out, void* stream) {
    if (!g || !out || K < 1 || P < 1) return pfail("bad argument to qb_colsq_mean");
    const unsigned blocks = (unsigned)((P + 127) / 128);
    if (dtype == QB_F64) k_colsq_mean<double><<<blocks, 128, 0, (cudaStream_t)stream>>>((const double*)g, K, P, out);
    else k_colsq_mean<float><<<blocks, 128, 0, (cudaStream_t)stream>>>((const float*)g, K, P, out);
    QP_LAUNCHED();
    return 0;
}

// ---- quantiles over the leading (member) axis of y[M, n]: out[iq, i] = quantile_q(y[:, i]) with numpy's default
// linear interpolation.  A block owns CB consecutive columns: loads them coalesced, sorts every column with a bitonic
// network in shared memory (one warp per column at a time), interpolates.
struct QbQuantArgs { double q[8]; int nq; };
template <typename T>
__global__ void __launch_bounds__(128) k_quantiles(const T* y, long long M, long long n, int Mp, int CB, QbQuantArgs qa, T* out) {
    extern __shared__ __align__(16) unsigned char qsm[];
    T* s = reinterpret_cast<T*>(qsm);                       // [CB][Mp]
    const long long c0 = (long long)blockIdx.x * CB;
    for (long long e = threadIdx.x; e < (long long)Mp * CB; e += blockDim.x) {
        const int c = (int)(e % CB);
        const long long m = e / CB;
        T v = (T)INFINITY;                                  // padding sorts to the end
        if (m < M && c0 + c < n) v = y[m * n + c0 + c];
        s[(long long)c * Mp + m] = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int c = warp; c < CB; c += nw) {
        T* col = s + (long long)c * Mp;
        for (int k = 2; k <= Mp; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = lane; i < Mp; i += 32) {
                    const int l = i ^ j;
                    if (l > i) {
                        const T a = col[i], b = col[l];
                        const bool up = (i & k) == 0;
                        if ((a > b) == up) { col[i] = b; col[l] = a; }
                    }
                }
                __syncwarp();
            }
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < qa.nq * CB; e += blockDim.x) {
        const int c = e % CB, iq = e / CB;
        if (c0 + c >= n) continue;
        const T* col = s + (long long)c * Mp;
        const double h = (double)(M - 1) * qa.q[iq];
        long long lo = (long long)floor(h);
        if (lo < 0) lo = 0;
        if (lo > M - 1) lo = M - 1;
        const long long hi = lo + 1 < M ? lo + 1 : M - 1;
        const double t = h - (double)lo, a = (double)col[lo], b = (double)col[hi];
        const double v = t < 0.5 ? a + (b - a) * t : b - (b - a) * (1.0 - t);     // numpy's _lerp
        out[(long long)iq * n + c0 + c] = (T)v;
    }
}
extern "C" int qb_quantiles(int dtype, const void* y, int64_t M, int64_t n, const double* q_host, int nq, void* out, void* stream) {
    if (!y || !out || !q_host || M < 1 || n < 1 || nq < 1 || nq > 8) return pfail("bad argument to qb_quantiles (1 <= nq <= 8)");
    int Mp = 1;
    while (Mp < M) Mp <<= 1;
    const int es = dtype == QB_F64 ? 8 : 4;
    const long long col_bytes = (long long)Mp * es;
    if (col_bytes > 200 * 1024) return pfail("qb_quantiles: too many members for one shared-memory column (M <= 25600 fp64 / 51200 fp32)");
    int CB = (int)((96 * 1024) / col_bytes);
    if (CB > 16) CB = 16;
    if (CB < 1) CB = 1;
    QbQuantArgs qa;
    qa.nq = nq;
    for (int i = 0; i < 8; ++i) qa.q[i] = i < nq ? q_host[i] : 0.0;
    for (int i = 0; i < nq; ++i) if (!(qa.q[i] >= 0.0 && qa.q[i] <= 1.0)) return pfail("qb_quantiles: q must be in [0, 1]");
    const size_t smem = (size_t)col_bytes * CB;
    const unsigned blocks = (unsigned)((n + CB - 1) / CB);
    if (dtype == QB_F64) {
        QP_CUDA(cudaFuncSetAttribute(k_quantiles<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_quantiles<double><<<blocks, 128, smem, (cudaStream_t)stream>>>((const double*)y, M, n, Mp, CB, qa, (double*)out);
    } else {
        QP_CUDA(cudaFuncSetAttribute(k_quantiles<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_quantiles<float><<<blocks, 128, smem, (cudaStream_t)stream>>>((const float*)y, M, n, Mp, CB, qa, (float*)out);
    }
    QP_LAUNCHED();
    return 0;
}

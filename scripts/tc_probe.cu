// Development probe for the tcgen05 path: checks the TMEM round trip and one 128x64x64 kind::tf32 MMA with A in
// TMEM and B in shared memory (K-major, no swizzle) against a host product.  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tc_probe scripts/tc_probe.cu && build/tc_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (int it = 0; it < (1 << 22); it++) {
        uint32_t ok;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

// variant bit0: swap LBO/SBO fields in the descriptor
__global__ void __launch_bounds__(128, 1) probe(const float* Ag, const float* Bg, float* Dg, float* RTg, int* status,
                                                int K, int N, int variant) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    float* Bs = (float*)smem;
    const uint32_t lbo = 128, sbo = 128u * (K / 4);
    for (int e = tid; e < N * K; e += 128) {
        int n = e / K, k = e % K;
        uint32_t off = (n >> 3) * sbo + (k >> 2) * lbo + (n & 7) * 16 + (k & 3) * 4;
        Bs[off / 4] = Bg[e];
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // generic-proxy smem writes must be visible to the async proxy (tensor core reads B through it)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    // A: row tid, K columns at [0, K)
    for (int c = 0; c < K; c += 8) {
        uint32_t r[8];
        for (int i = 0; i < 8; i++) r[i] = __float_as_uint(Ag[tid * K + c + i]);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     :: "r"(lane_base + c), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    // round trip check
    for (int c = 0; c < K; c += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(lane_base + c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; i++) RTg[tid * K + c + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // instruction descriptor: D=F32, A=B=TF32, K-major both, N, M=128
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        uint32_t f_lbo = (variant & 1) ? sbo : lbo, f_sbo = (variant & 1) ? lbo : sbo;
        for (int s = 0; s < K / 8; s++) {
            uint32_t addr = smem_u32(Bs) + s * 2 * lbo;
            uint64_t desc = (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((f_lbo >> 4) & 0x3FFF) << 16) |
                            ((uint64_t)((f_sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
            uint32_t acc = s > 0;
            asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p; }"
                         :: "r"(tbase + 128u), "r"(tbase + (uint32_t)(s * 8)), "l"(desc), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    bool ok = mbar_wait(smem_u32(&bar), 0);
    if (!ok) { if (tid == 0) *status = 1; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok) {
        for (int c = 0; c < N; c += 8) {
            uint32_t r[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(lane_base + 128u + c) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 8; i++) Dg[tid * N + c + i] = __uint_as_float(r[i]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0 && ok)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(256u) : "memory");
}

int main() {
    const int M = 128;
    int shapes[3][2] = {{64, 64}, {32, 16}, {16, 48}};          // {K, N}
    for (int sh = 0; sh < 3; sh++) {
        int K = shapes[sh][0], N = shapes[sh][1];
        float *A = (float*)malloc(M * K * 4), *B = (float*)malloc(N * K * 4), *D = (float*)malloc(M * N * 4), *RT = (float*)malloc(M * K * 4);
        srand(7 + sh);
        for (int i = 0; i < M * K; i++) A[i] = (float)(rand() % 9 - 4) * 0.25f;
        for (int i = 0; i < N * K; i++) B[i] = (float)(rand() % 9 - 4) * 0.5f;
        float *dA, *dB, *dD, *dRT; int* dS;
        CK(cudaMalloc(&dA, M * K * 4)); CK(cudaMalloc(&dB, N * K * 4)); CK(cudaMalloc(&dD, M * N * 4)); CK(cudaMalloc(&dRT, M * K * 4)); CK(cudaMalloc(&dS, 4));
        CK(cudaMemcpy(dA, A, M * K * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B, N * K * 4, cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        for (int variant = 0; variant < 2; variant++) {
            CK(cudaMemset(dD, 0, M * N * 4)); CK(cudaMemset(dS, 0, 4));
            probe<<<1, 128, 64 * 1024>>>(dA, dB, dD, dRT, dS, K, N, variant);
            CK(cudaGetLastError());
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("K=%d N=%d variant %d: launch failed: %s\n", K, N, variant, cudaGetErrorString(e)); return 1; }
            int st; CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(D, dD, M * N * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(RT, dRT, M * K * 4, cudaMemcpyDeviceToHost));
            double rt_err = 0, err = 0; int bad = 0;
            for (int i = 0; i < M * K; i++) rt_err = fmax(rt_err, fabs(RT[i] - A[i]));
            for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) {
                double ref = 0; for (int k = 0; k < K; k++) ref += (double)A[m * K + k] * B[n * K + k];
                double d = fabs(ref - D[m * N + n]); if (d > 1e-5) bad++; err = fmax(err, d);
            }
            printf("K=%d N=%d variant %d: status %d, tmem round-trip err %.3g, mma max err %.3g, bad %d / %d\n", K, N, variant, st, rt_err, err, bad, M * N);
            if (bad && variant == 0) {
                printf("  D[0][0..7] got:"); for (int n = 0; n < 8; n++) printf(" %g", D[n]);
                printf("\n  expected     :"); for (int n = 0; n < 8; n++) { double ref = 0; for (int k = 0; k < K; k++) ref += (double)A[k] * B[n * K + k]; printf(" %g", ref); }
                printf("\n");
            }
        }
        cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dRT); cudaFree(dS);
        free(A); free(B); free(D); free(RT);
    }
    return 0;
}

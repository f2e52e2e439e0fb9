// Development probe #4: cycle cost of the tcgen05.mma kind::tf32 shapes the gradient kernel issues (clock64 around a batch
// of MMAs + commit + mbarrier wait, one block on one SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tc_probe4 scripts/tc_probe4.cu && build/tc_probe4
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (int it = 0; it < (1 << 24); it++) {
        uint32_t ok;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
// mode 0: A in tensor memory (M=128), B K-major smem.  mode 1: A, B MN-major (layout 1) smem.  mode 2: A MN-major, B K-major (stride 144)
// Descriptors are base + compile-time constants (fully unrolled), as in the kernels: the loop measures the hardware, not
// the address arithmetic of the issuing thread.
template <int MODE, int M, int N, int NMMA, int ND = 1, int NW = 1>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 160 * 1024 / 4; e += 128) ((float*)smem)[e] = 0.001f * (e & 1023);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"((uint32_t)NW) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base_s;
    uint32_t phase = 0;
    long long best = 1LL << 60, best_issue = 0;
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (MODE >= 1) idesc |= 1u << 15;
    if (MODE == 1) idesc |= 1u << 16;
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 80 * 1024);
    const uint32_t a_lo = ((a0 >> 4) & 0x3FFF) | ((512u >> 4) << 16), a_hi = (1024u >> 4) | (1u << 14) | (1u << 29);
    const uint32_t bk_lo = ((b0 >> 4) & 0x3FFF) | ((128u >> 4) << 16), bk_hi = ((128u * 16) >> 4) | (1u << 14);
    const uint32_t bm_lo = ((b0 >> 4) & 0x3FFF) | ((512u >> 4) << 16), bm_hi = a_hi;
    const uint32_t bx_lo = ((b0 >> 4) & 0x3FFF) | ((144u >> 4) << 16), bx_hi = (4608u >> 4) | (1u << 14);
    for (int r = 0; r < reps; ++r) {
        long long t0 = 0, t1 = 0;
        __syncthreads();
        if ((tid & 31) == 0 && warp < NW) {
            t0 = clock64();
#pragma unroll
            for (int s = 0; s < NMMA / NW; s++) {
                const int ks = s % 16;
                if (MODE == 0) {
                    asm volatile("{ .reg .pred p; .reg .b64 db; setp.ne.b32 p, %5, 0; mov.b64 db, {%2, %3}; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p; }"
                                 :: "r"(tbase + 256u + (uint32_t)((s % ND) * 64) + (uint32_t)(warp * 64)), "r"(tbase + (uint32_t)((ks % 8) * 8)), "r"(bk_lo + (ks % 8) * 16), "r"(bk_hi), "r"(idesc), "r"(1u) : "memory");
                } else {
                    const uint32_t blo = MODE == 1 ? bm_lo + ks * 128 : bx_lo + ks * 18, bhi = MODE == 1 ? bm_hi : bx_hi;
                    asm volatile("{ .reg .pred p; .reg .b64 da, db; setp.ne.b32 p, %6, 0; mov.b64 da, {%1, %2}; mov.b64 db, {%3, %4}; tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p; }"
                                 :: "r"(tbase + 256u + (uint32_t)((s % ND) * 64) + (uint32_t)(warp * 64)), "r"(a_lo + ks * 128), "r"(a_hi), "r"(blo), "r"(bhi), "r"(idesc), "r"(1u) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
            t1 = clock64();
        }
        mbar_wait(smem_u32(&bar), phase);
        phase ^= 1;
        if (tid == 0) {
            const long long t2 = clock64();
            if (t2 - t0 < best) { best = t2 - t0; best_issue = t1 - t0; }
        }
        __syncthreads();
    }
    if (tid == 0) { out[0] = best; out[1] = best_issue; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(512u) : "memory");
}
template <int MODE, int M, int N, int NMMA, int ND = 1, int NW = 1>
int run(long long* d, const char* what) {
    CK(cudaFuncSetAttribute(probe<MODE, M, N, NMMA, ND, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    probe<MODE, M, N, NMMA, ND, NW><<<1, 128, 200 * 1024>>>(d, 20);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    long long h[2]; CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    printf("%-52s ND=%d NW=%d n=%3d: total %6lld cycles (%.1f / MMA), issue %lld (%.1f / MMA)\n", what, ND, NW, NMMA, h[0], (double)h[0] / NMMA, h[1], (double)h[1] / NMMA);
    return 0;
}
int main() {
    long long* d; CK(cudaMalloc(&d, 16));
    run<0, 128, 64, 24>(d, "FWD/BWD: A tmem, B K-major, M128 N64");
    run<0, 128, 64, 96, 1, 2>(d, "FWD/BWD, 2 issuing warps");
    run<0, 128, 64, 96, 1, 4>(d, "FWD/BWD, 4 issuing warps");
    run<1, 64, 64, 96, 1, 2>(d, "DW1, 2 issuing warps");
    run<1, 64, 64, 96, 1, 4>(d, "DW1, 4 issuing warps");
    run<2, 64, 8, 96, 1, 4>(d, "DB1/DW0 N8, 4 issuing warps");
    run<0, 128, 64, 96, 2>(d, "FWD/BWD, 2 accumulators interleaved");
    run<0, 128, 64, 96, 3>(d, "FWD/BWD, 3 accumulators interleaved");
    run<0, 128, 64, 96, 4>(d, "FWD/BWD, 4 accumulators interleaved");
    run<1, 64, 64, 96, 2>(d, "DW1, 2 accumulators interleaved");
    run<1, 64, 64, 96, 4>(d, "DW1, 4 accumulators interleaved");
    run<2, 64, 8, 96, 4>(d, "DB1/DW0 N8, 4 accumulators interleaved");
    run<1, 64, 32, 96, 4>(d, "A,B MN-major M64 N32, 4 accumulators");
    run<0, 128, 64, 96>(d, "same x96");
    run<0, 128, 32, 48>(d, "A tmem M128 N32");
    run<1, 64, 64, 48>(d, "DW1: A,B MN-major smem, M64 N64");
    run<1, 64, 64, 96>(d, "same x96");
    run<1, 128, 64, 48>(d, "A,B MN-major smem, M128 N64");
    run<1, 64, 32, 48>(d, "A,B MN-major M64 N32");
    run<1, 64, 72, 48>(d, "A,B MN-major M64 N72");
    run<1, 64, 8, 48>(d, "A,B MN-major M64 N8");
    run<2, 64, 8, 48>(d, "DB1/DW0: A MN-major, B K-major(144), M64 N8");
    run<2, 64, 8, 96>(d, "same x96");
    run<2, 64, 16, 48>(d, "A MN-major, B K-major(144), M64 N16");
    run<2, 128, 16, 48>(d, "A MN-major, B K-major(144), M128 N16");
    return 0;
}

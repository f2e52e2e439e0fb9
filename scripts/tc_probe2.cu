// Development probe for the tensor-core GRADIENT path: tcgen05.mma kind::tf32 with BOTH operands in shared memory,
// each either K-major or MN-major (no swizzle), M = 128 or 64, checked against a host product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tc_probe2 scripts/tc_probe2.cu && build/tc_probe2
// What it establishes (used by quinn_b200/csrc/qb_tcg.cuh):
//   * the shared-memory descriptor fields of an MN-major operand: SBO = stride between groups of 4 MN elements,
//     LBO = stride between groups of 8 K elements (K-major: LBO = stride between 4-element K chunks, SBO = stride
//     between groups of 8 rows);
//   * one buffer whose 128-byte core matrices hold (8 points) x (4 units) serves as the K-major operand
//     [points x units] AND as the MN-major operand [units x points];
//   * where the rows of an M = 64 accumulator live in tensor memory: lane (m % 16) + 32 * (m / 16).
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

// operand X[R][K] (logical, row-major in global) -> shared memory, 128-byte core matrices
//   K-major : core (r/8, k/4) holds 8 rows x 4 k      byte = (r/8)*SR + (k/4)*128 + (r%8)*16 + (k%4)*4, SR = 128*K/4
//   MN-major: core (k/8, r/4) holds 8 k    x 4 rows   byte = (k/8)*SK + (r/4)*128 + (k%8)*16 + (r%4)*4, SK = 128*R/4
__device__ void fill(float* S, const float* G, int R, int K, int mn_major) {
    for (int e = threadIdx.x; e < R * K; e += blockDim.x) {
        const int r = e / K, k = e % K;
        uint32_t off;
        if (!mn_major) off = (r >> 3) * (128u * (K / 4)) + (k >> 2) * 128u + (r & 7) * 16u + (k & 3) * 4u;
        else off = (k >> 3) * (128u * (R / 4)) + (r >> 2) * 128u + (k & 7) * 16u + (r & 3) * 4u;
        S[off / 4] = G[e];
    }
}

__device__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// variant bit0: swap LBO/SBO of A, bit1: swap LBO/SBO of B
__global__ void __launch_bounds__(128, 1) probe(const float* Ag, const float* Bg, float* Dg, int* status,
                                                int M, int N, int K, int a_mn, int b_mn, int variant) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* As = (float*)smem;
    float* Bs = (float*)(smem + 64 * 1024);
    fill(As, Ag, M, K, a_mn);
    fill(Bs, Bg, N, K, b_mn);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base_s;
    if (tid == 0) {
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        if (a_mn) idesc |= 1u << 15;
        if (b_mn) idesc |= 1u << 16;
        // K-major: LBO = 128 (k chunks), SBO = 128*K/4 (row groups), k-step = +256 bytes
        // MN-major: SBO = 128 (groups of 4 rows), LBO = 128*R/4 (groups of 8 k), k-step = +LBO
        uint32_t a_lbo = a_mn ? 128u * (M / 4) : 128u, a_sbo = a_mn ? 128u : 128u * (K / 4), a_step = a_mn ? 128u * (M / 4) : 256u;
        uint32_t b_lbo = b_mn ? 128u * (N / 4) : 128u, b_sbo = b_mn ? 128u : 128u * (K / 4), b_step = b_mn ? 128u * (N / 4) : 256u;
        if (variant & 1) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; }
        if (variant & 2) { uint32_t t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
        for (int s = 0; s < K / 8; s++) {
            const uint64_t da = make_desc(smem_u32(As) + s * a_step, a_lbo, a_sbo);
            const uint64_t db = make_desc(smem_u32(Bs) + s * b_step, b_lbo, b_sbo);
            uint32_t acc = s > 0;
            asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                         :: "r"(tbase), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    bool ok = mbar_wait(smem_u32(&bar), 0);
    if (!ok) { if (tid == 0) *status = 1; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok) {
        // dump ALL 128 lanes x N columns: Dg[lane][n]
        const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < N; c += 8) {
            uint32_t r[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(lane_base + c) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 8; i++) Dg[(warp * 32 + lane) * N + c + i] = __uint_as_float(r[i]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(256u) : "memory");
}

int main() {
    // {M, N, K, a_mn, b_mn}
    int cases[][5] = {
        {128, 64, 64, 0, 0},     // forward:  D[p][j] = A0[p][:] . W[j][:]        (both K-major)
        {128, 64, 64, 0, 1},     // backward: D[p][i] = Z[p][:] . W[:][i]         (B = W read MN-major)
        {64, 72, 128, 1, 1},     // dW:       D[j][i] = sum_p Z[p][j] A0[p][i]    (both MN-major, M = 64)
        {64, 8, 128, 1, 1},      // dW0:      narrow N
        {64, 40, 128, 1, 1},     // dW for 32-wide layers (M = 64 reads 32 rows of padding)
        {128, 32, 32, 0, 0},
        {128, 32, 32, 0, 1},
        {128, 16, 128, 1, 1},    // stacked M = 128 MN-major
    };
    const int ncases = sizeof(cases) / sizeof(cases[0]);
    for (int cs = 0; cs < ncases; cs++) {
        const int M = cases[cs][0], N = cases[cs][1], K = cases[cs][2], a_mn = cases[cs][3], b_mn = cases[cs][4];
        float *A = (float*)malloc(M * K * 4), *B = (float*)malloc(N * K * 4), *D = (float*)malloc(128 * N * 4);
        srand(7 + cs);
        for (int i = 0; i < M * K; i++) A[i] = (float)(rand() % 9 - 4) * 0.25f;
        for (int i = 0; i < N * K; i++) B[i] = (float)(rand() % 9 - 4) * 0.5f;
        float *dA, *dB, *dD; int* dS;
        CK(cudaMalloc(&dA, M * K * 4)); CK(cudaMalloc(&dB, N * K * 4)); CK(cudaMalloc(&dD, 128 * N * 4)); CK(cudaMalloc(&dS, 4));
        CK(cudaMemcpy(dA, A, M * K * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B, N * K * 4, cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        for (int variant = 0; variant < 4; variant++) {
            CK(cudaMemset(dD, 0, 128 * N * 4)); CK(cudaMemset(dS, 0, 4));
            probe<<<1, 128, 128 * 1024>>>(dA, dB, dD, dS, M, N, K, a_mn, b_mn, variant);
            CK(cudaGetLastError());
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("case %d variant %d: launch failed: %s\n", cs, variant, cudaGetErrorString(e)); return 1; }
            int st; CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(D, dD, 128 * N * 4, cudaMemcpyDeviceToHost));
            // two candidate row -> lane maps for M = 64: (m%16) + 32*(m/16) and identity
            for (int map = 0; map < (M == 64 ? 2 : 1); map++) {
                double err = 0; int bad = 0;
                for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) {
                    double ref = 0; for (int k = 0; k < K; k++) ref += (double)A[m * K + k] * B[n * K + k];
                    const int lane = (M == 64 && map == 0) ? (m % 16) + 32 * (m / 16) : m;
                    double d = fabs(ref - D[lane * N + n]); if (d > 1e-4) bad++; err = fmax(err, d);
                }
                printf("M=%d N=%d K=%d a_mn=%d b_mn=%d variant %d map %d: status %d, max err %.3g, bad %d / %d\n", M, N, K, a_mn, b_mn,
                       variant, map, st, err, bad, M * N);
            }
        }
        cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
        free(A); free(B); free(D);
    }
    return 0;
}

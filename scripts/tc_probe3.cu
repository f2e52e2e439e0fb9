// Development probe #3 for the tensor-core gradient path: hypotheses about shared-memory operand layouts of
// tcgen05.mma kind::tf32 (both operands in shared memory).  The host computes, per hypothesis, the byte offset of
// every logical element and the descriptor fields; the kernel only scatters and issues.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tc_probe3 scripts/tc_probe3.cu && build/tc_probe3
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <functional>
#include <string>
#include <algorithm>
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

struct Op { uint32_t lbo, sbo, step, layout_type, mn; };

__global__ void __launch_bounds__(128, 1) probe(const float* Ag, const int* Amap, int nA, const float* Bg, const int* Bmap, int nB,
                                                float* Dg, int* status, int M, int N, int K, Op oa, Op ob) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* As = (float*)smem;
    float* Bs = (float*)(smem + 96 * 1024);
    for (int e = tid; e < 48 * 1024; e += 128) { As[e] = 0.f; }
    for (int e = tid; e < 96 * 1024 / 4 / 4; e += 128) { Bs[e] = 0.f; }
    __syncthreads();
    for (int e = tid; e < nA; e += 128) As[Amap[e] / 4] = Ag[e];
    for (int e = tid; e < nB; e += 128) Bs[Bmap[e] / 4] = Bg[e];
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
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    // fill D with a sentinel so that "nothing written" is visible
    for (int c = 0; c < 256; c += 8) {
        const uint32_t s = __float_as_uint(-777.0f);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" :: "r"(lane_base + c), "r"(s) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        if (oa.mn) idesc |= 1u << 15;
        if (ob.mn) idesc |= 1u << 16;
        for (int s = 0; s < K / 8; s++) {
            const uint32_t aa = smem_u32(As) + s * oa.step, ab = smem_u32(Bs) + s * ob.step;
            const uint64_t da = (uint64_t)((aa >> 4) & 0x3FFF) | ((uint64_t)((oa.lbo >> 4) & 0x3FFF) << 16) |
                                ((uint64_t)((oa.sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)oa.layout_type << 61);
            const uint64_t db = (uint64_t)((ab >> 4) & 0x3FFF) | ((uint64_t)((ob.lbo >> 4) & 0x3FFF) << 16) |
                                ((uint64_t)((ob.sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)ob.layout_type << 61);
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

// ---------------------------------------------------------------------------------------------------------------------
// layout hypotheses: byte offset of logical element (r = MN index, k) of an operand with R rows, and descriptor fields
// ---------------------------------------------------------------------------------------------------------------------
struct Hyp { const char* name; Op op; std::function<int(int, int)> off; };

static Hyp hyp(const char* which, int R, int K) {
    Hyp h; h.name = which;
    std::string w(which);
    if (w == "K") {                 // K-major, no swizzle, LBO 128
        const int sbo = 128 * (K / 4);
        h.op = Op{128u, (uint32_t)sbo, 256u, 0u, 0u};
        h.off = [=](int r, int k) { return (r / 8) * sbo + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4; };
    } else if (w == "K144") {       // K-major, no swizzle, padded chunk stride (bank-conflict-free transposed writes)
        const int sbo = 144 * (K / 4);
        h.op = Op{144u, (uint32_t)sbo, 288u, 0u, 0u};
        h.off = [=](int r, int k) { return (r / 8) * sbo + (k / 4) * 144 + (r % 8) * 16 + (k % 4) * 4; };
    } else if (w == "MN32a" || w == "MN32b" || w == "MN32c" || w == "MN32d") {
        // MN-major, SWIZZLE_128B_BASE32B (layout type 1): atom = 4 k-rows x 128 bytes (32 MN elements); 32-byte unit index
        // XOR k-row.  a: desc lbo = MN-atom stride, sbo = K-atom stride; b: fields swapped; c/d: element-level XOR variant
        const int mn_atoms = (R + 31) / 32;
        const int katom = 512 * mn_atoms;         // K-atom stride (MN atoms of one K-atom adjacent)
        const bool sw = (w == "MN32b" || w == "MN32d"), elem = (w == "MN32c" || w == "MN32d");
        h.op = Op{sw ? (uint32_t)katom : 512u, sw ? 512u : (uint32_t)katom, (uint32_t)(2 * katom), 1u, 1u};
        h.off = [=](int r, int k) {
            const int base = (k / 4) * katom + (r / 32) * 512 + (k % 4) * 128;
            const int u = r % 32;
            if (!elem) return base + (((u / 8) ^ (k % 4)) * 32) + (u % 8) * 4;
            return base + (u / 4) * 16 + (((u % 4) ^ ((u / 4) & 3)) * 4);
        };
    } else if (w == "MN128a" || w == "MN128b") {
        // MN-major, SWIZZLE_128B (layout type 2): atom = 8 k-rows x 128 bytes; 16-byte chunk index XOR k-row
        const int mn_atoms = (R + 31) / 32;
        const int katom = 1024 * mn_atoms;
        const bool sw = (w == "MN128b");
        h.op = Op{sw ? (uint32_t)katom : 1024u, sw ? 1024u : (uint32_t)katom, (uint32_t)katom, 2u, 1u};
        h.off = [=](int r, int k) {
            const int u = r % 32;
            return (k / 8) * katom + (r / 32) * 1024 + (k % 8) * 128 + (((u / 4) ^ (k % 8)) * 16) + (u % 4) * 4;
        };
    } else if (w == "MN0a" || w == "MN0b") {
        // MN-major, no swizzle: core = 8 k x 4 MN
        const int kst = 128 * (R / 4);
        const bool sw = (w == "MN0b");
        h.op = Op{sw ? 128u : (uint32_t)kst, sw ? (uint32_t)kst : 128u, (uint32_t)kst, 0u, 1u};
        h.off = [=](int r, int k) { return (k / 8) * kst + (r / 4) * 128 + (k % 8) * 16 + (r % 4) * 4; };
    } else { printf("unknown hypothesis %s\n", which); exit(1); }
    return h;
}

int main() {
    struct Case { int M, N, K; const char* ha; const char* hb; };
    std::vector<Case> cases = {
        {128, 64, 64, "K", "K"},
        {64, 64, 128, "K144", "K144"},       // option T: transposed K-major operands with padded chunk stride
        {64, 8, 128, "K144", "K144"},
        {128, 64, 64, "K", "MN32a"}, {128, 64, 64, "K", "MN32b"}, {128, 64, 64, "K", "MN32c"}, {128, 64, 64, "K", "MN32d"},
        {128, 64, 64, "K", "MN128a"}, {128, 64, 64, "K", "MN128b"},
        {128, 64, 64, "K", "MN0a"}, {128, 64, 64, "K", "MN0b"},
        {64, 64, 128, "MN32a", "MN32a"}, {64, 64, 128, "MN32b", "MN32b"},
        {64, 64, 128, "MN32c", "MN32c"}, {64, 64, 128, "MN32d", "MN32d"},
        {64, 64, 128, "MN128a", "MN128a"}, {64, 64, 128, "MN128b", "MN128b"},
        {64, 8, 128, "MN32a", "K144"}, {64, 8, 128, "MN32b", "K144"},
        {64, 32, 128, "MN32a", "MN32a"}, {64, 32, 128, "MN32b", "MN32b"},
    };
    for (size_t cs = 0; cs < cases.size(); cs++) {
        const int M = cases[cs].M, N = cases[cs].N, K = cases[cs].K;
        Hyp ha = hyp(cases[cs].ha, M, K), hb = hyp(cases[cs].hb, N, K);
        std::vector<float> A(M * K), B(N * K), D(128 * N);
        std::vector<int> Am(M * K), Bm(N * K);
        srand(7 + (int)cs);
        int maxa = 0, maxb = 0;
        for (int i = 0; i < M * K; i++) { A[i] = (float)(rand() % 9 - 4) * 0.25f; Am[i] = ha.off(i / K, i % K); maxa = std::max(maxa, Am[i]); }
        for (int i = 0; i < N * K; i++) { B[i] = (float)(rand() % 9 - 4) * 0.5f; Bm[i] = hb.off(i / K, i % K); maxb = std::max(maxb, Bm[i]); }
        if (maxa >= 96 * 1024 || maxb >= 96 * 1024) { printf("case %zu: operand too large\n", cs); continue; }
        float *dA, *dB, *dD; int *dS, *dAm, *dBm;
        CK(cudaMalloc(&dA, M * K * 4)); CK(cudaMalloc(&dB, N * K * 4)); CK(cudaMalloc(&dD, 128 * N * 4)); CK(cudaMalloc(&dS, 4));
        CK(cudaMalloc(&dAm, M * K * 4)); CK(cudaMalloc(&dBm, N * K * 4));
        CK(cudaMemcpy(dA, A.data(), M * K * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), N * K * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dAm, Am.data(), M * K * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dBm, Bm.data(), N * K * 4, cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024));
        CK(cudaMemset(dD, 0, 128 * N * 4)); CK(cudaMemset(dS, 0, 4));
        probe<<<1, 128, 192 * 1024>>>(dA, dAm, M * K, dB, dBm, N * K, dD, dS, M, N, K, ha.op, hb.op);
        CK(cudaGetLastError());
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("case %zu: launch failed: %s\n", cs, cudaGetErrorString(e)); return 1; }
        int st; CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(D.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost));
        double err = 0; int bad = 0, untouched = 0;
        for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) {
            double ref = 0; for (int k = 0; k < K; k++) ref += (double)A[m * K + k] * B[n * K + k];
            const int lane = (M == 64) ? (m % 16) + 32 * (m / 16) : m;
            const float got = D[lane * N + n];
            if (got == -777.0f) untouched++;
            double d = fabs(ref - got); if (d > 1e-4) bad++; err = fmax(err, d);
        }
        printf("M=%d N=%d K=%d A=%s B=%s: status %d, max err %.3g, bad %d / %d, untouched %d  %s\n", M, N, K, ha.name, hb.name, st, err, bad,
               M * N, untouched, bad == 0 ? "OK" : "");
        cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS); cudaFree(dAm); cudaFree(dBm);
    }
    return 0;
}

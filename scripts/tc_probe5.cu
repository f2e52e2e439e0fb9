// Development probe #5 for the 128-wide tensor-core gradient path: shared-memory operand layouts of tcgen05.mma kind::f16
// (16-bit operands, both in shared memory, no swizzle, K-major and MN-major views of the SAME image; mixed fp16 / bf16).  The host computes, per hypothesis, the byte offset of
// every logical element and the descriptor fields; the kernel only scatters and issues.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tc_probe5 scripts/tc_probe5.cu && build/tc_probe5
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <functional>
#include <string>
#include <algorithm>
#include <cstring>
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

struct Op { uint32_t lbo, sbo, step, layout_type, mn, fmt; };

__global__ void __launch_bounds__(128, 1) probe(const uint16_t* Ag, const int* Amap, int nA, const uint16_t* Bg, const int* Bmap, int nB,
                                                float* Dg, int* status, int M, int N, int K, Op oa, Op ob) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint16_t* As = (uint16_t*)smem;
    uint16_t* Bs = (uint16_t*)(smem + 96 * 1024);
    for (int e = tid; e < 96 * 1024; e += 128) { As[e] = 0; }
    __syncthreads();
    for (int e = tid; e < nA; e += 128) As[Amap[e] / 2] = Ag[e];
    for (int e = tid; e < nB; e += 128) Bs[Bmap[e] / 2] = Bg[e];
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
        uint32_t idesc = (1u << 4) | (oa.fmt << 7) | (ob.fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        if (oa.mn) idesc |= 1u << 15;
        if (ob.mn) idesc |= 1u << 16;
        for (int s = 0; s < K / 16; s++) {
            const uint32_t aa = smem_u32(As) + s * oa.step, ab = smem_u32(Bs) + s * ob.step;
            const uint64_t da = (uint64_t)((aa >> 4) & 0x3FFF) | ((uint64_t)((oa.lbo >> 4) & 0x3FFF) << 16) |
                                ((uint64_t)((oa.sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)oa.layout_type << 61);
            const uint64_t db = (uint64_t)((ab >> 4) & 0x3FFF) | ((uint64_t)((ob.lbo >> 4) & 0x3FFF) << 16) |
                                ((uint64_t)((ob.sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)ob.layout_type << 61);
            uint32_t acc = s > 0;
            asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
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
// Images (byte offset of logical element):
//   P(R): "point-major" image of R-unit rows of 128 points: (p, u) at (u/8)*2048 + (p/8)*128 + (p%8)*16 + (u%8)*2
//   W   : weight image of tc3: (j, i) at (j/8)*2048 + (i/8)*128 + (j%8)*16 + (i%8)*2      (128 x 128)
// Views: K-major  (rows r = MN index, k): LBO = stride between 8-element k chunks, SBO = stride between 8-row groups
//        MN-major (r = MN index, k):      SBO = stride between 8-element MN chunks, LBO = stride between 8-row k groups  ("a")
//                                         or the two fields swapped ("b")
// ---------------------------------------------------------------------------------------------------------------------
struct Hyp { std::string name; Op op; std::function<int(int, int)> off; };
static int offP(int p, int u) { return (u / 8) * 2048 + (p / 8) * 128 + (p % 8) * 16 + (u % 8) * 2; }
static int offW(int j, int i) { return (j / 8) * 2048 + (i / 8) * 128 + (j % 8) * 16 + (i % 8) * 2; }
static Hyp hyp(const std::string& w, uint32_t fmt) {
    Hyp h; h.name = w + (fmt ? "/bf16" : "/f16");
    if (w == "P.K")        { h.op = Op{2048u, 128u, 4096u, 0u, 0u, fmt}; h.off = [](int r, int k) { return offP(r, k); }; }     // rows = points, k = units
    else if (w == "W.K")   { h.op = Op{128u, 2048u, 256u, 0u, 0u, fmt};  h.off = [](int r, int k) { return offW(r, k); }; }     // rows = j, k = i
    else if (w == "P.MNa") { h.op = Op{128u, 2048u, 256u, 0u, 1u, fmt};  h.off = [](int r, int k) { return offP(k, r); }; }     // rows = units, k = points
    else if (w == "P.MNb") { h.op = Op{2048u, 128u, 256u, 0u, 1u, fmt};  h.off = [](int r, int k) { return offP(k, r); }; }
    else if (w == "W.MNa") { h.op = Op{2048u, 128u, 4096u, 0u, 1u, fmt}; h.off = [](int r, int k) { return offW(k, r); }; }     // rows = i, k = j
    else if (w == "W.MNb") { h.op = Op{128u, 2048u, 4096u, 0u, 1u, fmt}; h.off = [](int r, int k) { return offW(k, r); }; }
    else { printf("unknown hypothesis %s\n", w.c_str()); exit(1); }
    return h;
}
static uint16_t enc(float v, uint32_t fmt) {          // values are multiples of 1/4 in [-2, 2]: exact in fp16 and bf16
    uint32_t b; memcpy(&b, &v, 4);
    if (fmt) return (uint16_t)(b >> 16);
    if (v == 0.0f) return 0;
    const uint32_t s = b >> 31, e = ((b >> 23) & 255) - 127 + 15, m = (b >> 13) & 1023;
    return (uint16_t)((s << 15) | (e << 10) | m);
}

int main() {
    struct Case { int M, N, K; const char* ha; uint32_t fa; const char* hb; uint32_t fb; };
    std::vector<Case> cases = {
        {128, 128, 128, "P.K", 0, "W.K", 0},                                   // FWD
        {128, 128, 128, "P.K", 0, "W.MNa", 0}, {128, 128, 128, "P.K", 0, "W.MNb", 0},   // BWD (B = W1 viewed MN-major)
        {128, 144, 128, "P.MNa", 0, "P.MNa", 0}, {128, 144, 128, "P.MNb", 0, "P.MNb", 0},   // DW1 + DB1
        {128, 16, 128, "P.MNa", 0, "P.MNa", 0},                               // DW0
        {128, 128, 128, "P.K", 1, "W.K", 1},                                   // all bf16
        {128, 128, 128, "P.K", 1, "W.MNa", 0}, {128, 144, 128, "P.MNa", 1, "P.MNa", 0},   // mixed: A bf16, B fp16
        {128, 16, 128, "P.MNa", 1, "P.MNa", 1},
    };
    for (size_t cs = 0; cs < cases.size(); cs++) {
        const int M = cases[cs].M, N = cases[cs].N, K = cases[cs].K;
        Hyp ha = hyp(cases[cs].ha, cases[cs].fa), hb = hyp(cases[cs].hb, cases[cs].fb);
        std::vector<float> A(M * K), B(N * K), D(128 * N);
        std::vector<uint16_t> Ah(M * K), Bh(N * K);
        std::vector<int> Am(M * K), Bm(N * K);
        srand(7 + (int)cs);
        int maxa = 0, maxb = 0;
        for (int i = 0; i < M * K; i++) { A[i] = (float)(rand() % 9 - 4) * 0.25f; Ah[i] = enc(A[i], cases[cs].fa); Am[i] = ha.off(i / K, i % K); maxa = std::max(maxa, Am[i]); }
        for (int i = 0; i < N * K; i++) { B[i] = (float)(rand() % 9 - 4) * 0.5f; Bh[i] = enc(B[i], cases[cs].fb); Bm[i] = hb.off(i / K, i % K); maxb = std::max(maxb, Bm[i]); }
        if (maxa >= 96 * 1024 || maxb >= 96 * 1024) { printf("case %zu: operand too large\n", cs); continue; }
        uint16_t *dA, *dB; float* dD; int *dS, *dAm, *dBm;
        CK(cudaMalloc(&dA, M * K * 2)); CK(cudaMalloc(&dB, N * K * 2)); CK(cudaMalloc(&dD, 128 * N * 4)); CK(cudaMalloc(&dS, 4));
        CK(cudaMalloc(&dAm, M * K * 4)); CK(cudaMalloc(&dBm, N * K * 4));
        CK(cudaMemcpy(dA, Ah.data(), M * K * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, Bh.data(), N * K * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dAm, Am.data(), M * K * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dBm, Bm.data(), N * K * 4, cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024));
        CK(cudaMemset(dD, 0, 128 * N * 4)); CK(cudaMemset(dS, 0, 4));
        probe<<<1, 128, 192 * 1024>>>(dA, dAm, M * K, dB, dBm, N * K, dD, dS, M, N, K, ha.op, hb.op);
        CK(cudaGetLastError());
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("case %zu (A=%s B=%s): launch failed: %s\n", cs, ha.name.c_str(), hb.name.c_str(), cudaGetErrorString(e)); return 1; }
        int st; CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(D.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost));
        double err = 0; int bad = 0, untouched = 0;
        for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) {
            double ref = 0; for (int k = 0; k < K; k++) ref += (double)A[m * K + k] * B[n * K + k];
            const float got = D[m * N + n];
            if (got == -777.0f) untouched++;
            double d = fabs(ref - got); if (d > 1e-4) bad++; err = fmax(err, d);
        }
        printf("M=%d N=%d K=%d A=%s B=%s: status %d, max err %.3g, bad %d / %d, untouched %d  %s\n", M, N, K, ha.name.c_str(), hb.name.c_str(), st, err, bad,
               M * N, untouched, bad == 0 ? "OK" : "");
        cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS); cudaFree(dAm); cudaFree(dBm);
    }
    return 0;
}

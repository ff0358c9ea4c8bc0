// Standalone validation of the tcgen05 / TMEM building blocks used by the fused marcher's MLP
// (csrc/k4_march_tc.cuh): one CTA evaluates a 3-layer MLP (K1 -> 128 -> 128 -> 3, ReLU) for a
// 128-row batch with
//   L1: tcgen05.mma kind::f16, A and B from shared memory (canonical K-major, no swizzle),
//   L2/L3: A from TENSOR MEMORY (the fp16-packed ReLU output of the previous layer, written with
//   tcgen05.st over the accumulator it was read from), B from shared memory,
// and compares with a CPU evaluation that applies the same fp16 roundings.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int M = 128, K1 = 48, W = 128, N3 = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// canonical K-major INTERLEAVE layout: (row r, 16-byte k-chunk kc) -> byte offset
__host__ __device__ inline int canon_off(int r, int kc, int kchunks) { return (r >> 3) * (kchunks * 128) + kc * 128 + (r & 7) * 16; }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int kchunks) {
    const uint64_t lbo = 128 >> 4, sbo = (uint64_t)(kchunks * 128) >> 4;
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);   // F32 accum, F16 x F16, K-major A/B
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n}\n"
                 :: "r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// weights arrive pre-packed in the canonical layout (host does it), X as [M][K1] fp16
__global__ void __launch_bounds__(128) mlp_kernel(const __half* X, const unsigned char* w1p, const unsigned char* w2p,
                                                  const unsigned char* w3p, const float* b1, const float* b2, const float* b3,
                                                  float* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;                              // 128 x 48 halves = 12288 B
    unsigned char* sW1 = sA + M * K1 * 2;                  // 128 x 48
    unsigned char* sW2 = sW1 + W * K1 * 2;                 // 128 x 128
    unsigned char* sW3 = sW2 + W * W * 2;                  // 16 x 128
    uint64_t* bar = reinterpret_cast<uint64_t*>(sW3 + N3 * W * 2);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;

    // stage weights (plain copies here; the marcher uses one TMA bulk copy)
    for (int i = tid; i < W * K1 * 2 / 16; i += 128) reinterpret_cast<uint4*>(sW1)[i] = reinterpret_cast<const uint4*>(w1p)[i];
    for (int i = tid; i < W * W * 2 / 16; i += 128) reinterpret_cast<uint4*>(sW2)[i] = reinterpret_cast<const uint4*>(w2p)[i];
    for (int i = tid; i < N3 * W * 2 / 16; i += 128) reinterpret_cast<uint4*>(sW3)[i] = reinterpret_cast<const uint4*>(w3p)[i];
    // A tile: thread r writes its row in the canonical layout
    for (int kc = 0; kc < K1 / 8; ++kc)
        *reinterpret_cast<uint4*>(sA + canon_off(tid, kc, K1 / 8)) = *reinterpret_cast<const uint4*>(X + tid * K1 + kc * 8);
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // generic smem writes -> async proxy (tensor core)
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tbase = *tmem_slot;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    const uint32_t D1 = 0, D2 = 128, D3 = 192;
    uint32_t phase = 0;

    // ---- layer 1 (SS) ----
    if (tid == 0) {
        const uint32_t idesc = make_idesc(M, W);
        for (int s = 0; s < K1 / 16; ++s)
            mma_ss(tbase + D1, make_desc(smem_u32(sA) + s * 256, K1 / 8), make_desc(smem_u32(sW1) + s * 256, K1 / 8), idesc, s > 0);
        umma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    for (int c = 0; c < 4; ++c) {                         // ReLU(D1 + b1) -> fp16 pairs, written over D1's first 64 columns
        uint32_t v[32], h[16];
        tmem_ld32(lane_base + D1 + c * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        for (int j = 0; j < 16; ++j)
            h[j] = pack_h2(fmaxf(__uint_as_float(v[2 * j]) + b1[c * 32 + 2 * j], 0.f), fmaxf(__uint_as_float(v[2 * j + 1]) + b1[c * 32 + 2 * j + 1], 0.f));
        tmem_st16(lane_base + D1 + c * 16, h);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

    // ---- layer 2 (TS: A = H1 in TMEM) ----
    if (tid == 0) {
        const uint32_t idesc = make_idesc(M, W);
        for (int s = 0; s < W / 16; ++s)
            mma_ts(tbase + D2, tbase + D1 + s * 8, make_desc(smem_u32(sW2) + s * 256, W / 8), idesc, s > 0);
        umma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    for (int c = 0; c < 4; ++c) {
        uint32_t v[32], h[16];
        tmem_ld32(lane_base + D2 + c * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        for (int j = 0; j < 16; ++j)
            h[j] = pack_h2(fmaxf(__uint_as_float(v[2 * j]) + b2[c * 32 + 2 * j], 0.f), fmaxf(__uint_as_float(v[2 * j + 1]) + b2[c * 32 + 2 * j + 1], 0.f));
        tmem_st16(lane_base + D2 + c * 16, h);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

    // ---- layer 3 (TS, N = 16) ----
    if (tid == 0) {
        const uint32_t idesc = make_idesc(M, N3);
        for (int s = 0; s < W / 16; ++s)
            mma_ts(tbase + D3, tbase + D2 + s * 8, make_desc(smem_u32(sW3) + s * 256, W / 8), idesc, s > 0);
        umma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    {
        uint32_t v[4];
        tmem_ld4(lane_base + D3, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        for (int j = 0; j < 3; ++j) out[tid * 3 + j] = __uint_as_float(v[j]) + b3[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tbase), "r"(256u) : "memory");
}

static float h2f(__half h) { return __half2float(h); }

int main() {
    srand(1);
    auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    std::vector<float> X(M * K1), W1(W * K1), W2(W * W), W3(3 * W), b1(W), b2(W), b3(3);
    for (auto& v : X) v = rnd();
    for (int k = 39; k < K1; ++k) for (int r = 0; r < M; ++r) X[r * K1 + k] = 0.f;
    for (auto& v : W1) v = rnd() * 0.16f;
    for (auto& v : W2) v = rnd() * 0.09f;
    for (auto& v : W3) v = rnd() * 0.09f;
    for (auto& v : b1) v = rnd() * 0.1f;
    for (auto& v : b2) v = rnd() * 0.1f;
    for (auto& v : b3) v = rnd() * 0.1f;
    std::vector<__half> Xh(M * K1);
    for (int i = 0; i < M * K1; ++i) Xh[i] = __float2half_rn(X[i]);
    auto pack = [](const std::vector<float>& Wm, int n_out, int npad, int k) {
        std::vector<unsigned char> p((size_t)npad * k * 2, 0);
        for (int n = 0; n < n_out; ++n)
            for (int kk = 0; kk < k; ++kk) {
                __half h = __float2half_rn(Wm[(size_t)n * k + kk]);
                memcpy(&p[canon_off(n, kk / 8, k / 8) + (kk % 8) * 2], &h, 2);
            }
        return p;
    };
    auto w1p = pack(W1, W, W, K1), w2p = pack(W2, W, W, W), w3p = pack(W3, 3, N3, W);
    // CPU reference with the same fp16 roundings of operands, fp32 accumulate
    std::vector<float> ref(M * 3);
    for (int r = 0; r < M; ++r) {
        float h1[W], h2[W];
        for (int n = 0; n < W; ++n) {
            float a = 0.f;
            for (int k = 0; k < K1; ++k) a += h2f(Xh[r * K1 + k]) * h2f(__float2half_rn(W1[n * K1 + k]));
            h1[n] = h2f(__float2half_rn(fmaxf(a + b1[n], 0.f)));
        }
        for (int n = 0; n < W; ++n) {
            float a = 0.f;
            for (int k = 0; k < W; ++k) a += h1[k] * h2f(__float2half_rn(W2[n * W + k]));
            h2[n] = h2f(__float2half_rn(fmaxf(a + b2[n], 0.f)));
        }
        for (int n = 0; n < 3; ++n) {
            float a = 0.f;
            for (int k = 0; k < W; ++k) a += h2[k] * h2f(__float2half_rn(W3[n * W + k]));
            ref[r * 3 + n] = a + b3[n];
        }
    }
    __half* dX; unsigned char *d1, *d2, *d3; float *db1, *db2, *db3, *dout;
    CK(cudaMalloc(&dX, Xh.size() * 2)); CK(cudaMalloc(&d1, w1p.size())); CK(cudaMalloc(&d2, w2p.size())); CK(cudaMalloc(&d3, w3p.size()));
    CK(cudaMalloc(&db1, W * 4)); CK(cudaMalloc(&db2, W * 4)); CK(cudaMalloc(&db3, 16)); CK(cudaMalloc(&dout, M * 3 * 4));
    CK(cudaMemcpy(dX, Xh.data(), Xh.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d1, w1p.data(), w1p.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d2, w2p.data(), w2p.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d3, w3p.data(), w3p.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db1, b1.data(), W * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db2, b2.data(), W * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db3, b3.data(), 12, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0, M * 3 * 4));
    const int smem = M * K1 * 2 + W * K1 * 2 + W * W * 2 + N3 * W * 2 + 64;
    CK(cudaFuncSetAttribute(mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mlp_kernel<<<1, 128, smem>>>(dX, d1, d2, d3, db1, db2, db3, dout);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> out(M * 3);
    CK(cudaMemcpy(out.data(), dout, M * 3 * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int i = 0; i < M * 3; ++i) { maxerr = fmax(maxerr, fabs(out[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); }
    printf("{\"micro\": \"umma_mlp_test\", \"max_abs_err\": %.3e, \"max_abs_ref\": %.3e, \"out0\": [%f, %f, %f], \"ref0\": [%f, %f, %f], \"ok\": %s}\n",
           maxerr, maxref, out[0], out[1], out[2], ref[0], ref[1], ref[2], maxerr < 2e-3 ? "true" : "false");
    return 0;
}

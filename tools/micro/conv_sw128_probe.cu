// Probe for the next decoder-conv step (profiles/r1_conv_ws_ncu.md: the high-resolution convs are bound by the
// copy engine's 16-byte-row rate): can the im2col-free 3x3 conv keep its "taps = descriptor start addresses"
// trick when the halo is staged with 128-BYTE rows (64 channels per pixel, SWIZZLE_128B K-major) by ONE TMA
// tensor copy?  A tap shifts the A operand's start address by a multiple of 128 B, i.e. off the 1024-byte
// swizzle atom, so the descriptor's base-offset field (bits 49-51) comes into play.  This program stages an
// 18 x 34 x 64ch halo and 9 x [64][64] weight tiles with TMA (SWIZZLE_128B), runs the 9 x 4 x 4 (taps x M tiles x
// k16 steps) tcgen05 MMAs for three descriptor conventions, and compares each with a CPU convolution:
//   mode 0: base_offset = (start_address >> 7) & 7     (what the PTX ISA text suggests)
//   mode 1: base_offset = 0                             (swizzle taken from absolute address bits)
//   mode 2: as 0, and the k16 step folded into the start address only (same as 0; sanity duplicate)
// Integer-valued inputs make the comparison exact.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int HY = 18, HX = 34, C = 64, N = 64, TM = 4;
constexpr int ROW = 128;                       // bytes per pixel (64 fp16)
constexpr int A_BYTES = HY * HX * ROW;         // 78336
constexpr int A_PAD = (A_BYTES + 1023) / 1024 * 1024;
constexpr int B_TAP = N * ROW;                 // 8192
constexpr int SMEM = A_PAD + 9 * B_TAP + 1024 + 64;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
           ((uint64_t)(base_off & 7) << 49) | (2ull << 61);          // layout_type 2 = SWIZZLE_128B
}
__device__ __forceinline__ uint32_t idesc(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n"
                 :: "r"(d), "l"(a), "l"(b), "r"(id), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, float* out, int mode) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* A = smem;
    unsigned char* B = smem + A_PAD;
    uint64_t* bar = reinterpret_cast<uint64_t*>(B + 9 * B_TAP);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(s32(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(s32(bar + 1)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" :: "r"(s32(tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tbase = *tslot;
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(s32(bar)), "r"((uint32_t)(A_BYTES + 9 * B_TAP)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                     :: "r"(s32(A)), "l"(reinterpret_cast<uint64_t>(&ta)), "r"(s32(bar)), "r"(0), "r"(0), "r"(0) : "memory");
        for (int t = 0; t < 9; ++t)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
                         :: "r"(s32(B + t * B_TAP)), "l"(reinterpret_cast<uint64_t>(&tb)), "r"(s32(bar)), "r"(0), "r"(t * N) : "memory");
    }
    mbar_wait(bar, 0);
    uint32_t leader;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(leader));
    if (warp == 0 && leader) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t id = idesc(128, N);
        for (int m = 0; m < TM; ++m)
            for (int t = 0; t < 9; ++t) {
                const int dy = t / 3, dx = t % 3;
                for (int ks = 0; ks < C / 16; ++ks) {
                    const uint32_t a_addr = s32(A) + (uint32_t)((dy * HX + dx + m * 8) * ROW) + ks * 32;
                    const uint32_t b_addr = s32(B) + t * B_TAP + ks * 32;
                    const uint32_t a_bo = (mode == 1) ? 0u : ((a_addr >> 7) & 7);
                    mma_ss(tbase + m * N, desc_sw128(a_addr, HX * ROW, a_bo), desc_sw128(b_addr, 1024, 0), id, (t | ks) != 0);
                }
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(s32(bar + 1)) : "memory");
    }
    mbar_wait(bar + 1, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tl = tbase + ((uint32_t)(warp * 32) << 16);
    for (int m = 0; m < TM; ++m)
        for (int c16 = 0; c16 < N / 16; ++c16) {
            uint32_t v[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(tl + m * N + c16 * 16));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            for (int j = 0; j < 16; ++j) out[((size_t)m * 128 + tid) * N + c16 * 16 + j] = __uint_as_float(v[j]);
        }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" :: "r"(tbase) : "memory");
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    CK(cudaSetDevice(0));
    CK(cudaFree(0));
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    if (!fp || q != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    encode_fn enc = reinterpret_cast<encode_fn>(fp);

    std::vector<__half> X((size_t)HY * HX * C), Wt((size_t)9 * N * C);
    uint32_t s = 12345;
    auto rnd = [&](int mod, int off) { s = s * 1664525u + 1013904223u; return (float)((int)((s >> 16) % mod) - off); };
    for (auto& v : X) v = __float2half(rnd(5, 2));
    for (auto& v : Wt) v = __float2half(rnd(3, 1));
    __half *dX, *dW; float* dO;
    CK(cudaMalloc(&dX, X.size() * 2)); CK(cudaMalloc(&dW, Wt.size() * 2)); CK(cudaMalloc(&dO, (size_t)TM * 128 * N * 4));
    CK(cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dW, Wt.data(), Wt.size() * 2, cudaMemcpyHostToDevice));

    alignas(64) CUtensorMap ta, tb;
    {
        const cuuint64_t gdim[3] = {C, HX, HY}, gstr[2] = {C * 2, (cuuint64_t)HX * C * 2};
        const cuuint32_t box[3] = {C, HX, HY}, es[3] = {1, 1, 1};
        if (enc(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, dX, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode A failed\n"); return 1; }
    }
    {
        const cuuint64_t gdim[2] = {C, 9 * N}, gstr[1] = {C * 2};
        const cuuint32_t box[2] = {C, N}, es[2] = {1, 1};
        if (enc(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dW, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode B failed\n"); return 1; }
    }
    std::vector<float> ref((size_t)TM * 128 * N, 0.f), got(ref.size());
    for (int m = 0; m < TM; ++m)
        for (int r = 0; r < 128; ++r) {
            const int y = r >> 3, x = m * 8 + (r & 7);
            for (int n = 0; n < N; ++n) {
                float acc = 0.f;
                for (int t = 0; t < 9; ++t)
                    for (int k = 0; k < C; ++k)
                        acc += __half2float(X[((size_t)(y + t / 3) * HX + x + t % 3) * C + k]) * __half2float(Wt[((size_t)t * N + n) * C + k]);
                ref[((size_t)m * 128 + r) * N + n] = acc;
            }
        }
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    for (int mode = 0; mode < 3; ++mode) {
        CK(cudaMemset(dO, 0, got.size() * 4));
        probe<<<1, 128, SMEM>>>(ta, tb, dO, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
        CK(cudaMemcpy(got.data(), dO, got.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0; size_t bad = 0;
        for (size_t i = 0; i < got.size(); ++i) { const double d = fabs((double)got[i] - ref[i]); if (d > maxerr) maxerr = d; if (d > 0.5) ++bad; }
        printf("mode %d (%s): max |err| = %g, mismatching = %zu / %zu  -> %s\n", mode,
               mode == 1 ? "base_offset 0" : "base_offset (addr>>7)&7", maxerr, bad, got.size(), bad == 0 ? "EXACT" : "WRONG");
    }
    return 0;
}

// k4_sr.cu -- the VC-Decoder (SFTNet, lib/sr_esrnet.py:400-465) on sm_100a.
//
// Replaces the reference's 229 cuDNN convolutions + torch.cat copies + F.interpolate + leaky_relu
// launches per tile with
//
//   conv3x3_ws_kernel<N>   every 3x3 convolution as an im2col-free implicit GEMM on tcgen05, one persistent
//       warp-specialised CTA per SM (TMA producer thread / elected MMA lane / two epilogue warpgroups, 5-6 stage
//       shared-memory ring, two TMEM accumulator sets).  The (16+2)x(32+2) halo of a 16-channel slice is staged
//       as two 8-channel planes [plane][hy][hx] x 16 B -- which IS the canonical K-major UMMA layout (core
//       matrix = 8 horizontally adjacent pixels, SBO = halo row pitch, LBO = plane pitch) -- so the taps are
//       shared-memory descriptors that differ only in their start address; weights are pre-packed per
//       (slice, tap) as [N][16] K-major tiles; fp32 accumulators live in TMEM for the whole K loop.
//       Dense-block concatenation is free (a conv reads channels [0,Cin) of the block's NHWC buffer and
//       writes its growth channels behind them); "nearest x2 then conv" runs in its sub-pixel form (four
//       phase convolutions over the source grid, 2x2 taps, phase-combined weights); bias / LeakyReLU /
//       residual scaling / trunk update are the epilogue.  See the comment at the kernel.
//   conv3x3_tc_kernel<N>   the first version (one 16x8 tile per CTA, 2-stage cp.async), kept as the A/B
//       reference of the persistent kernel (K4_CONV_V1=1);
//   sft_tc_kernel<COUT>    SFTLayer (lib/sr_esrnet.py:112-123): both 1x1-conv branches as two chained K=32
//       tcgen05 GEMMs per 128-pixel tile (A of the second from TMEM) + modulation as the epilogue;
//       sft_kernel is its fp32 one-thread-per-pixel predecessor (K4_SFT_FP32=1);
//   condnet_kernel         CondNet (lib/sr_esrnet.py:440-444), one thread per pixel, fp32.
//
// Numerics: conv operands are fp16 (weights and activations), accumulation fp32 in TMEM, the
// residual trunk and all CondNet math stay fp32.  (The reference itself runs these convs with
// TF32 operands: torch.backends.cudnn.allow_tf32 defaults to True.)
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "k4_internal.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// tcgen05 plumbing (same encodings as k4_march_tc.cu; cute/arch/mma_sm100_desc.hpp)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sr_s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sr_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t sr_idesc(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void sr_mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
// One lane of a CONVERGED warp.  Issuing tcgen05.mma under `if (elect)` instead of `if (tid == 0)` matters:
// ptxas knows an elected region is single-threaded, keeps the descriptors in uniform registers and emits
// back-to-back UTCHMMA; under a plain divergent branch it wraps every MMA in an ELECT / R2UR / BRA.U.ANY
// uniformisation loop (~12 extra instructions per MMA on the one thread that paces the tensor pipe).
__device__ __forceinline__ bool sr_elect_one() {
    uint32_t p;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ void sr_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(sr_s32(bar)) : "memory");
}
__device__ __forceinline__ void sr_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(sr_s32(bar)), "r"(count));
}
__device__ __forceinline__ void sr_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(sr_s32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void sr_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
// fp32 pair -> fp16x2, round to nearest, SATURATING (|x| > 65504 -> +-65504 instead of inf): activations of a
// trained checkpoint may exceed the fp16 range; a clamped operand degrades gracefully, an inf poisons the frame.
__device__ __forceinline__ __half2 sr_h2sat(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;\n" : "=r"(d) : "f"(hi), "f"(lo));
    return *reinterpret_cast<__half2*>(&d);
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(sr_s32(dst_smem)), "l"(src), "r"(src_bytes) : "memory");
}

constexpr int SR_TM = 1;                             // M tiles (16x8 pixels each) per CTA, side by side in x, sharing one
                                                     // weight stage.  Measured: TM=4 (1 CTA/SM) 34.5 ms/frame vs TM=1
                                                     // (2 CTAs/SM) 30.3 ms -- occupancy beats weight reuse here
constexpr int SR_TY = 16, SR_TX = 8 * SR_TM;         // output region of a CTA (pixels)
constexpr int SR_HY = SR_TY + 2, SR_HX = SR_TX + 2;  // halo
constexpr int SR_ROW = SR_HX * 16;                   // 160 B (TM=1): one halo row of one 8-channel plane
constexpr int SR_PLANE = SR_HY * SR_ROW;             // 2880 B (TM=1)
constexpr int SR_CK = 16;                            // input channels per K slice (one k16 MMA step per tap): small
                                                     // stages => 4 CTAs/SM whose load / MMA / epilogue phases overlap
constexpr int SR_A_STAGE = (SR_CK / 8) * SR_PLANE;   // 11520 B (TM=1)

enum { SRM_STORE_F16 = 0,      // dst_h[c0..c0+N) = act(acc + b)
       SRM_TRUNK = 1,          // dst_f = (acc + b) * scale + add_f                  (conv5: x5*0.2 + x)
       SRM_ADD_STORE_F16 = 2,  // dst_h = fp16((acc + b) + add_f)                    (conv_body + feat)
       SRM_STORE_F32F16 = 3,   // dst_f = acc + b (fp32, 64 ch) and dst_h = fp16     (conv_first: trunk + feat)
       SRM_OUT_NCHW = 4,       // out_nchw[c][y][x] = acc + b, c < n_valid          (conv_last)
       // conv3x3_ws_kernel only: the SFT layer(s) that follow the convolution run in its epilogue (see epilogue_sft)
       SRM_STORE_F16_SFT = 5,  // dst_h = fp16(sft(lrelu(acc + b)))                          (conv4 -> sft1)
       SRM_TRUNK_SFT = 6,      // t = (acc + b) * scale [+ add_f]; dst_f (, dst_f2) = t; dst_h2 = fp16(sft(t))   (conv_first, conv5 of rdb1/2 -> next sft0)
       SRM_TRUNK_SFT2 = 7 };   // t as above; u = sft(t) * scale2 + add_f2; dst_f = u; dst_h2 = fp16(sft2(u))    (conv5 of rdb3 -> RRDB tail -> next sft0 / sftbody)

struct ConvParams {
    const __half* src; int src_cstride; int src_c0; int cin;       // cin: multiple of 32 (zero padded weights beyond the real Cin)
    int H, W; int upsample;                                        // output size; upsample: src is (H/2) x (W/2), nearest
    const unsigned char* wpack; const float* bias;
    int mode; float lrelu; float scale;
    __half* dst_h; int dst_cstride; int dst_c0;
    float* dst_f; const float* add_f;                              // fp32 [P,64]
    float* out_nchw; int n_valid;
    int tiles_x;
    int use_tma;                                                   // conv3x3_ws_kernel: halo by TMA tensor copies (not with upsample)
    // conv3x3_ws_kernel only -- sub-pixel form of "nearest x2 + 3x3 conv": the convolution iterates over the SOURCE
    // grid with `ntaps` (4) of the nine halo offsets and phase-combined weights, and writes output pixel
    // (y*out_s + out_oy, x*out_s + out_ox) of an (H*out_s) x (W*out_s) destination.  Defaults: 9 taps, out_s = 1.
    int ntaps; unsigned char tap_dy[9], tap_dx[9];
    int out_s, out_oy, out_ox;
    // rows [y_lo, y_hi) of the iteration grid computed by this launch (all kernels; y_hi == 0 means all H rows): a unit
    // of the multi-GPU decoder only needs, per layer, the rows inside the remaining receptive field of its kept block
    int y_lo, y_hi;
    float* dst_f2;                                                 // SRM_STORE_F32F16: second fp32 copy (the initial trunk)
    // SRM_OUT_NCHW: output pixel (y, x) of channel c goes to out_nchw[c*out_ps + (y-crop_y0)*out_rs + (x-crop_x0)] when it lies
    // in the crop window [crop_y0,crop_y1) x [crop_x0,crop_x1) (the kept block of a tile, written straight into the frame)
    long long out_ps, out_rs; int crop_y0, crop_y1, crop_x0, crop_x1;
    // ... and at the same offset into out_extra[0 .. n_out_extra): the same window of the other ranks' frames (peer-mapped,
    // k4_srnet_forward_roi_peers) -- the multi-GPU exchange as NVLink stores of the kernel that produces the pixels
    int n_out_extra; float* out_extra[K4_MAX_PEERS - 1];
    // fused SFT epilogues (SRM_*_SFT*): fp16 condition map [P,32], the fragment-ordered operand blocks of the SFT layer(s)
    // (sft_frag_layout), the fp16 destination of the (last) modulated result, the RRDB input for the block tail
    const __half* cond16; const unsigned char* sftw; const unsigned char* sftw2; int sft_n;
    __half* dst_h2; int dst2_cstride, dst2_c0; const float* add_f2; float scale2;
};

__device__ __forceinline__ void sr_store_out(const ConvParams& p, long long idx, float v) {
    p.out_nchw[idx] = v;
#pragma unroll 1
    for (int e = 0; e < p.n_out_extra; ++e) p.out_extra[e][idx] = v;
}

// SFT operands for the in-epilogue mma.sync evaluation: fp16 weights W0 [64][32] (scale_conv0 rows then shift_conv0 rows),
// W1s [n][32], W1h [n][32] with 40-half row pitch (conflict-free fragment loads), then fp32 biases b0 [64], b1s + 1 [n], b1h [n].
constexpr int SFTF_PITCH = 40;
struct SftFrag { int off_w0, off_w1s, off_w1h, off_b0, off_b1s, off_b1h, total; };
__host__ __device__ inline SftFrag sft_frag_layout(int n) {
    SftFrag L; int o = 0;
    L.off_w0 = o; o += 64 * SFTF_PITCH * 2;
    L.off_w1s = o; o += n * SFTF_PITCH * 2;
    L.off_w1h = o; o += n * SFTF_PITCH * 2;
    L.off_b0 = o; o += 64 * 4;
    L.off_b1s = o; o += n * 4;
    L.off_b1h = o; o += n * 4;
    L.total = (o + 15) & ~15;
    return L;
}
constexpr int SFTF_MAX = 2 * (3 * 64 * SFTF_PITCH * 2 + 3 * 64 * 4);      // two 64-channel layers: 32,256 B

__device__ __forceinline__ void sr_hmma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t sr_pack_sat(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;\n" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// Hidden layer of one SFT layer for 16 pixels: cond fragments `ca` (m16 x k32) x W0 (64 outputs) -> LeakyReLU -> the A
// fragments of the two second GEMMs (scale branch: hidden 0..31, shift branch: hidden 32..63).  The accumulator fragment
// of two adjacent n8 blocks IS the A fragment of one k16 step (rows g / g+8, columns 2t..2t+1 / +8).
__device__ __forceinline__ void sft_hidden(const unsigned char* fw, const SftFrag& L, const uint32_t (&ca)[2][4], int g, int tq,
                                           uint32_t (&as)[2][4], uint32_t (&ah)[2][4]) {
    const __half* W0 = reinterpret_cast<const __half*>(fw + L.off_w0);
    const float* b0 = reinterpret_cast<const float*>(fw + L.off_b0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float c[4];
        c[0] = c[2] = b0[8 * j + 2 * tq]; c[1] = c[3] = b0[8 * j + 2 * tq + 1];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const __half* wr = W0 + (8 * j + g) * SFTF_PITCH + 16 * ks + 2 * tq;
            sr_hmma(c, ca[ks], *reinterpret_cast<const uint32_t*>(wr), *reinterpret_cast<const uint32_t*>(wr + 8));
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) c[e] = fmaxf(c[e], 0.2f * c[e]);
        uint32_t (&dst)[2][4] = (j < 4) ? as : ah;
        const int jj = j & 3;
        dst[jj >> 1][(jj & 1) * 2 + 0] = sr_pack_sat(c[0], c[1]);          // row g
        dst[jj >> 1][(jj & 1) * 2 + 1] = sr_pack_sat(c[2], c[3]);          // row g + 8
    }
}
// scale + 1 and shift of output channels 8j .. 8j+7 for the 16 pixels (accumulator fragment layout)
__device__ __forceinline__ void sft_scale_shift(const unsigned char* fw, const SftFrag& L, const uint32_t (&as)[2][4], const uint32_t (&ah)[2][4],
                                                int j, int g, int tq, float (&sc)[4], float (&sh)[4]) {
    const __half* W1s = reinterpret_cast<const __half*>(fw + L.off_w1s);
    const __half* W1h = reinterpret_cast<const __half*>(fw + L.off_w1h);
    const float* b1s = reinterpret_cast<const float*>(fw + L.off_b1s);
    const float* b1h = reinterpret_cast<const float*>(fw + L.off_b1h);
    sc[0] = sc[2] = b1s[8 * j + 2 * tq]; sc[1] = sc[3] = b1s[8 * j + 2 * tq + 1];
    sh[0] = sh[2] = b1h[8 * j + 2 * tq]; sh[1] = sh[3] = b1h[8 * j + 2 * tq + 1];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        const __half* ws = W1s + (8 * j + g) * SFTF_PITCH + 16 * ks + 2 * tq;
        const __half* wh = W1h + (8 * j + g) * SFTF_PITCH + 16 * ks + 2 * tq;
        sr_hmma(sc, as[ks], *reinterpret_cast<const uint32_t*>(ws), *reinterpret_cast<const uint32_t*>(ws + 8));
        sr_hmma(sh, ah[ks], *reinterpret_cast<const uint32_t*>(wh), *reinterpret_cast<const uint32_t*>(wh + 8));
    }
}
// 16 lanes x 8*NB columns of an fp32 accumulator in the mma.sync m16n8 accumulator-fragment layout: registers 4j..4j+3 of a
// thread = (row g, cols 8j+2t, +1), (row g+8, same cols) -- tools/micro/tmem_ld_shapes.cu checks this on the device
template <int NB> __device__ __forceinline__ void sr_ld_frag(uint32_t taddr, uint32_t (&v)[4 * NB]);
template <> __device__ __forceinline__ void sr_ld_frag<4>(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
}
template <> __device__ __forceinline__ void sr_ld_frag<8>(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr));
}

// Programmatic dependent launch: every decoder kernel is launched with programmaticStreamSerialization, so its CTAs are
// scheduled while the previous kernel drains; the set-up above this point (barrier init, TMEM allocation, constant
// weights) overlaps the predecessor's tail.  Nothing produced by an earlier kernel may be touched before the wait; the
// dependents are released only after it, so at most two consecutive kernels are ever in flight.
#define SR_PDL_SYNC() do { asm volatile("griddepcontrol.wait;\n" ::: "memory"); \
                           asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); } while (0)

template <int N>
__global__ void __launch_bounds__(128) conv3x3_tc_kernel(const __grid_constant__ ConvParams p) {
    constexpr int B_TAP = N * SR_CK * 2;               // bytes of one tap's [N][32] tile
    constexpr int B_STAGE = 9 * B_TAP;
    constexpr int STAGE = SR_A_STAGE + B_STAGE;
    constexpr int NACC = (N < 16) ? 16 : N;              // accumulator columns per M tile
    constexpr int TCOLS = (SR_TM * NACC < 32) ? 32 : SR_TM * NACC;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + 2 * STAGE);      // [2] MMA-done per stage
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + 2 * STAGE + 16);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x - ty * p.tiles_x;
    const int y0 = p.y_lo + ty * SR_TY, x0 = tx * SR_TX;

    if (tid == 0) {
        sr_mbar_init(mbar + 0, 1); sr_mbar_init(mbar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(sr_s32(tslot)), "r"((uint32_t)TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tbase = *tslot;
    SR_PDL_SYNC();

    const int sH = p.upsample ? (p.H >> 1) : p.H, sW = p.upsample ? (p.W >> 1) : p.W;
    const int nchunks = p.cin / SR_CK;

    auto load_stage = [&](int c, int st) {
        unsigned char* A = smem + st * STAGE;
        unsigned char* B = A + SR_A_STAGE;
        // halo: 180 pixels x 4 planes of 16 B
        for (int i = tid; i < SR_HY * SR_HX * (SR_CK / 8); i += 128) {
            const int plane = i % (SR_CK / 8), pix = i / (SR_CK / 8);
            const int hy = pix / SR_HX, hx = pix - hy * SR_HX;
            const int gy = y0 - 1 + hy, gx = x0 - 1 + hx;
            const bool in = (gy >= 0) & (gy < p.H) & (gx >= 0) & (gx < p.W);
            const int sy = p.upsample ? (gy >> 1) : gy, sx = p.upsample ? (gx >> 1) : gx;
            const __half* src = p.src;
            if (in) src = p.src + ((size_t)sy * sW + sx) * p.src_cstride + p.src_c0 + c * SR_CK + plane * 8;
            cp_async16(A + plane * SR_PLANE + hy * SR_ROW + hx * 16, src, in ? 16 : 0);
        }
        const unsigned char* wsrc = p.wpack + (size_t)c * B_STAGE;
        for (int i = tid; i < B_STAGE / 16; i += 128) cp_async16(B + i * 16, wsrc + i * 16, 16);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        (void)sH;
    };

    load_stage(0, 0);
    uint32_t ph[2] = {0, 0};
    for (int c = 0; c < nchunks; ++c) {
        const int st = c & 1;
        if (c + 1 < nchunks) {
            if (c >= 1) { sr_mbar_wait(mbar + (st ^ 1), ph[st ^ 1]); ph[st ^ 1] ^= 1; }   // MMAs of slice c-1 are done with that stage
            load_stage(c + 1, st ^ 1);
            asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t a0 = sr_s32(smem + st * STAGE), b0 = a0 + SR_A_STAGE;
            const uint32_t idesc = sr_idesc(128, NACC);
#pragma unroll 1
            for (int m = 0; m < SR_TM; ++m) {
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int dy = t / 3, dx = t - dy * 3;
#pragma unroll
                    for (int s = 0; s < SR_CK / 16; ++s) {
                        const uint64_t ad = sr_desc(a0 + (2 * s) * SR_PLANE + dy * SR_ROW + (m * 8 + dx) * 16, SR_PLANE, SR_ROW);
                        const uint64_t bd = sr_desc(b0 + t * B_TAP + s * 256, 128, (SR_CK / 8) * 128);
                        sr_mma_ss(tbase + m * NACC, ad, bd, idesc, (c | t | s) != 0);
                    }
                }
            }
            sr_commit(mbar + st);
        }
    }
    {   // all MMAs done: the last commit covers everything issued before it
        const int st = (nchunks - 1) & 1;
        sr_mbar_wait(mbar + st, ph[st]);
    }
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

    // ---------------- epilogue: thread r = y*8 + x owns one output pixel ----------------
    const uint32_t tl = tbase + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int m = 0; m < SR_TM; ++m) {
    const int py = y0 + (tid >> 3), px = x0 + m * 8 + (tid & 7);
    const bool inside = (py < p.y_hi) & (px < p.W);
    const size_t pix = (size_t)py * p.W + px;
#pragma unroll
    for (int c16 = 0; c16 < NACC / 16; ++c16) {
        uint32_t v[16];
        sr_ld16(tl + m * NACC + c16 * 16, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (!inside) continue;
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]) + __ldg(p.bias + c16 * 16 + j);
        if (p.mode == SRM_STORE_F16) {
            if (p.lrelu > 0.f) {
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] = o[j] > 0.f ? o[j] : o[j] * p.lrelu;
            }
            __half2 h[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) h[j] = sr_h2sat(o[2 * j], o[2 * j + 1]);
            uint4* d = reinterpret_cast<uint4*>(p.dst_h + pix * p.dst_cstride + p.dst_c0 + c16 * 16);
            d[0] = *reinterpret_cast<uint4*>(&h[0]);
            d[1] = *reinterpret_cast<uint4*>(&h[4]);
        } else if (p.mode == SRM_TRUNK) {
            const float4* a = reinterpret_cast<const float4*>(p.add_f + pix * 64 + c16 * 16);
            float4* d = reinterpret_cast<float4*>(p.dst_f + pix * 64 + c16 * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 r = a[q];
                d[q] = make_float4(o[4 * q] * p.scale + r.x, o[4 * q + 1] * p.scale + r.y, o[4 * q + 2] * p.scale + r.z, o[4 * q + 3] * p.scale + r.w);
            }
        } else if (p.mode == SRM_ADD_STORE_F16) {
            const float4* a = reinterpret_cast<const float4*>(p.add_f + pix * 64 + c16 * 16);
            __half2 h[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 r = a[q];
                h[2 * q] = sr_h2sat(o[4 * q] + r.x, o[4 * q + 1] + r.y);
                h[2 * q + 1] = sr_h2sat(o[4 * q + 2] + r.z, o[4 * q + 3] + r.w);
            }
            uint4* d = reinterpret_cast<uint4*>(p.dst_h + pix * p.dst_cstride + p.dst_c0 + c16 * 16);
            d[0] = *reinterpret_cast<uint4*>(&h[0]);
            d[1] = *reinterpret_cast<uint4*>(&h[4]);
        } else if (p.mode == SRM_STORE_F32F16) {
            float4* d = reinterpret_cast<float4*>(p.dst_f + pix * 64 + c16 * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q) d[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
            if (p.dst_f2) {
                float4* d2 = reinterpret_cast<float4*>(p.dst_f2 + pix * 64 + c16 * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) d2[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
            }
        } else {  // SRM_OUT_NCHW
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (c16 * 16 + j < p.n_valid && py >= p.crop_y0 && py < p.crop_y1 && px >= p.crop_x0 && px < p.crop_x1)
                    sr_store_out(p, (long long)(c16 * 16 + j) * p.out_ps + (long long)(py - p.crop_y0) * p.out_rs + (px - p.crop_x0), o[j]);
        }
    }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tbase), "r"((uint32_t)TCOLS) : "memory");
}

// ---------------------------------------------------------------------------------------------
// conv3x3_ws_kernel<N>: the same implicit GEMM as conv3x3_tc_kernel, restructured the Blackwell way --
// ONE persistent CTA per SM, warp specialised, fed through a deep shared-memory ring:
//   warps 4-7  producers: cp.async the (16+2)x(32+2) halo of a 16-channel slice (two 8-channel planes,
//              zero fill = the conv's padding, nearest-x2 folded into the addressing) plus that slice's
//              nine [N][16] weight tiles into ring stage g % NST; a stage is published (mbarrier FULL)
//              two stages behind the issue point, so every producer thread keeps three stages of
//              copies in flight;
//   warp 8     one elected thread issues 36 tcgen05.mma per stage (4 M tiles of 16x8 pixels x 9 taps,
//              K = 16) -- four M tiles share every weight tile, which cuts the L2->SM weight traffic
//              of the one-tile-per-CTA kernel by 4x -- and commits the stage back to the producers
//              (mbarrier EMPTY) and, after the last slice, the accumulator set to the epilogue;
//   warps 0-3  epilogue: thread = pixel, tcgen05.ld 16 columns at a time, bias / LeakyReLU / residual /
//              stores exactly as conv3x3_tc_kernel; TMEM holds TWO accumulator sets (2 x 4 x N
//              columns), so the epilogue of tile i overlaps the main loop of tile i+1.
// Tiles (16 rows x 32 columns of output pixels) are dealt round-robin to the CTAs.
// ---------------------------------------------------------------------------------------------
constexpr int WS_TM = 4;
constexpr int WS_TX = 8 * WS_TM, WS_HX = WS_TX + 2;            // 32 / 34
constexpr int WS_ROW = WS_HX * 16;                             // 544 B
constexpr int WS_PLANE = (SR_HY * WS_ROW + 127) / 128 * 128;   // 9856 B: 9792 B of pixels, padded so every plane is a 128-B aligned TMA destination
constexpr int WS_A_STAGE = (SR_CK / 8) * WS_PLANE;             // 19712 B
constexpr int WS_HALO = SR_HY * WS_HX * (SR_CK / 8);           // 1224 16-byte copies per stage
constexpr int WS_HALO_PER_THREAD = (WS_HALO + 127) / 128;      // 10
constexpr int WS_LAG = 2;
constexpr int WS_THREADS = 320;                               // warps 0-3 epilogue, 4-7 epilogue (TMA) / producers (cp.async), 8 MMA, 9 TMA producer

template <int N> struct WsCfg {
    static constexpr int B_TAP = N * SR_CK * 2;
    static constexpr int B_STAGE = 9 * B_TAP;
    static constexpr int STAGE = WS_A_STAGE + B_STAGE;
    static constexpr int NST = (N >= 64) ? 5 : 6;
    static constexpr int NACC = (N < 16) ? 16 : N;
    static constexpr int TCOLS = 2 * WS_TM * NACC;              // 512 / 256 / 128
    static constexpr int SMEM = NST * STAGE + 512;              // + barriers (256 B) + bias (256 B)
};

__device__ __forceinline__ void ws_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(sr_s32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Epilogue with the following SFT layer(s) folded in (SRM_*_SFT* modes).  The 36 SFT layers of the decoder are 0.2 % of
// its FLOPs but were 26 % of a tile's time as separate passes over the activations (profiles/r2_summary.md); here they
// cost no memory pass at all.  Per 16 pixels a warp evaluates the layer's two small MLPs with mma.sync (HMMA) straight
// from registers -- cond row fragments (fp16, [P,32]) x W0 -> LeakyReLU -> re-used as A fragments -> W1s / W1h -- and reads the
// convolution's accumulator from tensor memory in the SAME fragment layout (tcgen05.ld.16x256b), so scale, shift and the
// value they modulate meet in registers with no shuffle.  No tensor-memory scratch, no handshake with the MMA warp.
// Rows of a thread: pixels g and g+8 of the 16 (two image rows of the 16x8-pixel M tile), channels 8j+2t, +1 per block j.
// ---------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void epilogue_sft(const ConvParams& p, const unsigned char* fw, const float* sbias, uint32_t tacc0,
                                             int m_lo, int m_hi, int y0, int x0, int q, int lane) {
    constexpr int NB = N / 8;
    constexpr int NACC = N;
    const int g = lane >> 2, tq = lane & 3;
    const SftFrag L1 = sft_frag_layout(p.sft_n);
    const SftFrag L2 = sft_frag_layout(64);
    const unsigned char* fw2 = fw + L1.total;
    const bool two = (N == 64) && p.mode == SRM_TRUNK_SFT2;
    const bool trunk = p.mode != SRM_STORE_F16_SFT;
    const int nblk = 2 * (m_hi - m_lo);                          // 16-pixel blocks of this warp: (M tile, half)
    // Every global load of a block is issued before the block's arithmetic (the cond fragments of block b+1 even before
    // block b's): the epilogue is a chain of dependent steps on 8 warps, a load issued where it is used costs its full
    // latency (the first version of this function spent 2/3 of its time there).
    auto coords = [&](int blk, size_t& pixA, bool& inA, bool& inB) {
        const int m = m_lo + (blk >> 1), h = blk & 1;
        const int rA = 32 * q + 16 * h + g;                       // row of the M tile: pixel (rA >> 3, rA & 7); the second is rA + 8
        const int pyA = y0 + (rA >> 3), px = x0 + m * 8 + (rA & 7);
        inA = (pyA < p.y_hi) & (px < p.W); inB = (pyA + 1 < p.y_hi) & (px < p.W);
        pixA = (size_t)pyA * p.W + px;
    };
    auto load_cond = [&](int blk, uint32_t (&c)[2][4]) {
        size_t pixA; bool inA, inB;
        coords(blk, pixA, inA, inB);
        const __half* cA = p.cond16 + (inA ? pixA : 0) * 32 + 2 * tq;
        const __half* cB = p.cond16 + (inB ? pixA + p.W : 0) * 32 + 2 * tq;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            c[ks][0] = __ldg(reinterpret_cast<const unsigned int*>(cA + 16 * ks));
            c[ks][1] = __ldg(reinterpret_cast<const unsigned int*>(cB + 16 * ks));
            c[ks][2] = __ldg(reinterpret_cast<const unsigned int*>(cA + 16 * ks + 8));
            c[ks][3] = __ldg(reinterpret_cast<const unsigned int*>(cB + 16 * ks + 8));
        }
    };
    uint32_t cn[2][4];
    load_cond(0, cn);
#pragma unroll 1
    for (int blk = 0; blk < nblk; ++blk) {
        size_t pixA; bool inA, inB;
        coords(blk, pixA, inA, inB);
        const size_t pixB = pixA + p.W;
        uint32_t ca[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int e = 0; e < 4; ++e) ca[ks][e] = cn[ks][e];
        if (blk + 1 < nblk) load_cond(blk + 1, cn);
        // residual rows (fp32 trunk, and the RRDB input for the block tail) travel through a rolling window of PF channel
        // blocks: block j+PF is requested when block j is consumed -- enough bytes in flight to cover the memory latency
        // without holding the whole 256-byte rows in registers (which spills: 168 registers per thread is the limit here)
        constexpr int PF = (NB < 4) ? NB : 4;
        float2 rA[PF], rB[PF], xA[PF], xB[PF];
        auto load_res = [&](int j, int slot) {
            rA[slot] = rB[slot] = make_float2(0.f, 0.f);
            if (p.add_f) {
                if (inA) rA[slot] = __ldg(reinterpret_cast<const float2*>(p.add_f + pixA * 64 + 8 * j + 2 * tq));
                if (inB) rB[slot] = __ldg(reinterpret_cast<const float2*>(p.add_f + pixB * 64 + 8 * j + 2 * tq));
            }
            if (two) {
                xA[slot] = xB[slot] = make_float2(0.f, 0.f);
                if (inA) xA[slot] = __ldg(reinterpret_cast<const float2*>(p.add_f2 + pixA * 64 + 8 * j + 2 * tq));
                if (inB) xB[slot] = __ldg(reinterpret_cast<const float2*>(p.add_f2 + pixB * 64 + 8 * j + 2 * tq));
            }
        };
        if (trunk) {
#pragma unroll
            for (int j = 0; j < PF; ++j) load_res(j, j);
        }
        const uint32_t tblk = tacc0 + (uint32_t)((m_lo + (blk >> 1)) * NACC) + ((uint32_t)(16 * (blk & 1)) << 16);
        uint32_t as[2][4], ah[2][4], as2[2][4], ah2[2][4];
        sft_hidden(fw, L1, ca, g, tq, as, ah);
        if (two) sft_hidden(fw2, L2, ca, g, tq, as2, ah2);
#pragma unroll 1
        for (int jg = 0; jg < NB; jg += PF)
#pragma unroll
        for (int jj = 0; jj < PF; ++jj) {
            const int j = jg + jj;
            uint32_t acc[4];
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];\n"
                         : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]) : "r"(tblk + 8 * j));
            float sc[4], sh[4];
            sft_scale_shift(fw, L1, as, ah, j, g, tq, sc, sh);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            const int ch = 8 * j + 2 * tq;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = __uint_as_float(acc[e]) + sbias[ch + (e & 1)];
            if (!trunk) {                       // SRM_STORE_F16_SFT
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float a = p.lrelu > 0.f ? fmaxf(v[e], p.lrelu * v[e]) : v[e]; v[e] = fmaf(a, sc[e], sh[e]); }
                if (inA) *reinterpret_cast<uint32_t*>(p.dst_h + pixA * p.dst_cstride + p.dst_c0 + ch) = sr_pack_sat(v[0], v[1]);
                if (inB) *reinterpret_cast<uint32_t*>(p.dst_h + pixB * p.dst_cstride + p.dst_c0 + ch) = sr_pack_sat(v[2], v[3]);
                continue;
            }
            const int sl = jj;
            float t[4] = {fmaf(v[0], p.scale, rA[sl].x), fmaf(v[1], p.scale, rA[sl].y), fmaf(v[2], p.scale, rB[sl].x), fmaf(v[3], p.scale, rB[sl].y)};
            const float2 xa = xA[sl], xb = xB[sl];
            if (j + PF < NB) load_res(j + PF, sl);
            if (!two) {                         // SRM_TRUNK_SFT
                if (inA) {
                    *reinterpret_cast<float2*>(p.dst_f + pixA * 64 + ch) = make_float2(t[0], t[1]);
                    if (p.dst_f2) *reinterpret_cast<float2*>(p.dst_f2 + pixA * 64 + ch) = make_float2(t[0], t[1]);
                }
                if (inB) {
                    *reinterpret_cast<float2*>(p.dst_f + pixB * 64 + ch) = make_float2(t[2], t[3]);
                    if (p.dst_f2) *reinterpret_cast<float2*>(p.dst_f2 + pixB * 64 + ch) = make_float2(t[2], t[3]);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) t[e] = fmaf(t[e], sc[e], sh[e]);
            } else {                            // SRM_TRUNK_SFT2: u = sft(t) * scale2 + x_in -> trunk; then the next layer on u
                t[0] = fmaf(fmaf(t[0], sc[0], sh[0]), p.scale2, xa.x); t[1] = fmaf(fmaf(t[1], sc[1], sh[1]), p.scale2, xa.y);
                t[2] = fmaf(fmaf(t[2], sc[2], sh[2]), p.scale2, xb.x); t[3] = fmaf(fmaf(t[3], sc[3], sh[3]), p.scale2, xb.y);
                if (inA) *reinterpret_cast<float2*>(p.dst_f + pixA * 64 + ch) = make_float2(t[0], t[1]);
                if (inB) *reinterpret_cast<float2*>(p.dst_f + pixB * 64 + ch) = make_float2(t[2], t[3]);
                sft_scale_shift(fw2, L2, as2, ah2, j, g, tq, sc, sh);
#pragma unroll
                for (int e = 0; e < 4; ++e) t[e] = fmaf(t[e], sc[e], sh[e]);
            }
            if (inA) *reinterpret_cast<uint32_t*>(p.dst_h2 + pixA * p.dst2_cstride + p.dst2_c0 + ch) = sr_pack_sat(t[0], t[1]);
            if (inB) *reinterpret_cast<uint32_t*>(p.dst_h2 + pixB * p.dst2_cstride + p.dst2_c0 + ch) = sr_pack_sat(t[2], t[3]);
        }
    }
}

template <int N>
__global__ void __launch_bounds__(WS_THREADS, 1) conv3x3_ws_kernel(const __grid_constant__ ConvParams p, const __grid_constant__ CUtensorMap tmap) {
    using C = WsCfg<N>;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::NST * C::STAGE);     // [NST] count 128 (producers)
    uint64_t* empty = full + C::NST;                                             // [NST] count 1 (tcgen05.commit)
    uint64_t* acc_full = empty + C::NST;                                         // [2]   count 1 (tcgen05.commit)
    uint64_t* acc_empty = acc_full + 2;                                          // [2]   count 128 (epilogue)
    uint32_t* tslot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* sbias = reinterpret_cast<float*>(smem + C::NST * C::STAGE + 256);    // [64]
    unsigned char* sftw_s = smem + C::NST * C::STAGE + 512;                      // SFT operand block(s) of the fused epilogue modes
    const int tid = threadIdx.x, warp = tid >> 5;
    const int n_epi = p.use_tma ? 256 : 128;                                     // epilogue threads (8 or 4 warps)
    if (tid < 64) sbias[tid] = p.bias[tid];
    if (p.mode >= SRM_STORE_F16_SFT) {                                           // constants: may be staged before the PDL wait
        const int n1 = sft_frag_layout(p.sft_n).total;
        for (int i = tid; i < n1 / 16; i += WS_THREADS) reinterpret_cast<uint4*>(sftw_s)[i] = __ldg(reinterpret_cast<const uint4*>(p.sftw) + i);
        if (p.mode == SRM_TRUNK_SFT2) {
            const int n2 = sft_frag_layout(64).total;
            for (int i = tid; i < n2 / 16; i += WS_THREADS) reinterpret_cast<uint4*>(sftw_s + n1)[i] = __ldg(reinterpret_cast<const uint4*>(p.sftw2) + i);
        }
    }
    const int tiles_x = (p.W + WS_TX - 1) / WS_TX, tiles_y = (p.y_hi - p.y_lo + SR_TY - 1) / SR_TY;
    const int n_tiles = tiles_x * tiles_y;
    const int nchunks = p.cin / SR_CK;

    if (tid == 0) {
        for (int i = 0; i < C::NST; ++i) { sr_mbar_init(full + i, p.use_tma ? 1 : 128); sr_mbar_init(empty + i, 1); }
        sr_mbar_init(acc_full + 0, 1); sr_mbar_init(acc_full + 1, 1);
        sr_mbar_init(acc_empty + 0, n_epi); sr_mbar_init(acc_empty + 1, n_epi);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(sr_s32(tslot)), "r"((uint32_t)C::TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tbase = *tslot;
    SR_PDL_SYNC();

    if (warp == 9) {
        // ------------------------------ producer (TMA): one thread, three bulk copies per stage ------------------------------
        // The activation buffer is a 3-D tensor (channel, x, y); a box of (8 channels, 34, 18) lands in shared
        // memory as [hy][hx] x 16 B = one plane of the UMMA layout; out-of-range coordinates (the -1 halo ring,
        // the tile overhang) are zero-filled by the copy engine = the convolution's zero padding.
        if (p.use_tma && tid == 288) {
            unsigned g = 0;
            const uint64_t tm = reinterpret_cast<uint64_t>(&tmap);
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
                const int y0 = p.y_lo + ty * SR_TY - 1, x0 = tx * WS_TX - 1;
                for (int c = 0; c < nchunks; ++c, ++g) {
                    const unsigned slot = g % C::NST;
                    if (g >= (unsigned)C::NST) sr_mbar_wait(empty + slot, ((g / C::NST) - 1) & 1);
                    const uint32_t bar = sr_s32(full + slot);
                    const uint32_t A = sr_s32(smem + slot * C::STAGE);
                    const uint32_t bbytes = (uint32_t)p.ntaps * C::B_TAP;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar), "r"((uint32_t)((SR_CK / 8) * SR_HY * WS_ROW) + bbytes) : "memory");
#pragma unroll
                    for (int pl = 0; pl < SR_CK / 8; ++pl)
                        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                                     :: "r"(A + pl * WS_PLANE), "l"(tm), "r"(bar), "r"(c * SR_CK + pl * 8), "r"(x0), "r"(y0) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                                 :: "r"(A + WS_A_STAGE), "l"(p.wpack + (size_t)c * bbytes), "r"(bbytes), "r"(bar) : "memory");
                }
            }
        }
    } else if (warp >= 4 && warp < 8 && !p.use_tma) {
        // ------------------------------ producers (cp.async; nearest-x2 source addressing) ------------------------------
        const int pt = tid - 128;
        const int sW = p.upsample ? (p.W >> 1) : p.W;
        unsigned g = 0;
        auto publish = [&](unsigned gi) {
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            ws_arrive(full + gi % C::NST);
        };
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
            const int y0 = p.y_lo + ty * SR_TY, x0 = tx * WS_TX;
            long long soff[WS_HALO_PER_THREAD];               // element offset of this thread's halo copies (-1: zero fill)
#pragma unroll
            for (int j = 0; j < WS_HALO_PER_THREAD; ++j) {
                const int i = pt + j * 128;
                const int plane = i & 1, pix = i >> 1;
                const int hy = pix / WS_HX, hx = pix - hy * WS_HX;
                const int gy = y0 - 1 + hy, gx = x0 - 1 + hx;
                const bool in = (i < WS_HALO) & (gy >= 0) & (gy < p.H) & (gx >= 0) & (gx < p.W);
                const int sy = p.upsample ? (gy >> 1) : gy, sx = p.upsample ? (gx >> 1) : gx;
                soff[j] = in ? ((long long)sy * sW + sx) * p.src_cstride + p.src_c0 + plane * 8 : -1;
            }
            for (int c = 0; c < nchunks; ++c, ++g) {
                const unsigned slot = g % C::NST;
                if (g >= (unsigned)C::NST) sr_mbar_wait(empty + slot, ((g / C::NST) - 1) & 1);
                unsigned char* A = smem + slot * C::STAGE;
#pragma unroll
                for (int j = 0; j < WS_HALO_PER_THREAD; ++j) {
                    const int i = pt + j * 128;
                    if (i < WS_HALO) {
                        const int plane = i & 1, pix = i >> 1;
                        const int hy = pix / WS_HX, hx = pix - hy * WS_HX;
                        const bool in = soff[j] >= 0;
                        cp_async16(A + plane * WS_PLANE + hy * WS_ROW + hx * 16, in ? p.src + soff[j] + c * SR_CK : p.src, in ? 16 : 0);
                    }
                }
                const int bbytes = p.ntaps * C::B_TAP;
                const unsigned char* wsrc = p.wpack + (size_t)c * bbytes;
                unsigned char* B = A + WS_A_STAGE;
                for (int i = pt; i < bbytes / 16; i += 128) cp_async16(B + i * 16, wsrc + i * 16, 16);
                asm volatile("cp.async.commit_group;\n" ::: "memory");
                if (g >= (unsigned)WS_LAG) {
                    asm volatile("cp.async.wait_group %0;\n" :: "n"(WS_LAG) : "memory");
                    publish(g - WS_LAG);
                }
            }
        }
        // drain: the last WS_LAG stages
        if (g >= 2) { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); publish(g - 2); }
        if (g >= 1) { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); publish(g - 1); }
    } else if (warp == 8) {
        // ------------------------------ MMA issuer (one elected lane) ------------------------------
        if (sr_elect_one()) {
            const uint32_t idesc = sr_idesc(128, C::NACC);
            unsigned g = 0, it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const unsigned a = it & 1;
                if (it >= 2) sr_mbar_wait(acc_empty + a, ((it >> 1) - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t dacc = tbase + a * WS_TM * C::NACC;
                for (int c = 0; c < nchunks; ++c, ++g) {
                    const unsigned slot = g % C::NST;
                    sr_mbar_wait(full + slot, (g / C::NST) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint32_t a0 = sr_s32(smem) + slot * C::STAGE;
                    // descriptors differ from tap to tap only in the 14-bit start-address field: add constants
                    const uint64_t ad0 = sr_desc(a0, WS_PLANE, WS_ROW);
                    const uint64_t bd0 = sr_desc(a0 + WS_A_STAGE, 128, (SR_CK / 8) * 128);
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        if (t < p.ntaps) {
                            const uint32_t aoff = (uint32_t)(p.tap_dy[t] * WS_ROW + p.tap_dx[t] * 16) >> 4;
#pragma unroll
                            for (int m = 0; m < WS_TM; ++m)
                                sr_mma_ss(dacc + m * C::NACC, ad0 + (uint64_t)(aoff + m * 8), bd0 + (uint64_t)((t * C::B_TAP) >> 4),
                                          idesc, (c | t) != 0);
                        }
                    }
                    sr_commit(empty + slot);
                }
                sr_commit(acc_full + a);
            }
        }
    } else if (warp < 8) {
        // ------------------------------ epilogue ------------------------------
        // thread = pixel of an M tile (TMEM lane = 32*(warp%4) + lane).  With the TMA producer warps 4-7 are a
        // second epilogue warpgroup: group 0 takes M tiles 0,1 and group 1 tiles 2,3.  Per M tile all the
        // independent work is issued first -- the residual row (global, fp32) and every 16-column TMEM load --
        // and waited for once, so a tile's epilogue costs a few memory latencies, not one per 16 channels.
        constexpr int NCH = C::NACC / 16;
        const int et = tid & 127, egrp = tid >> 7;
        const int m_lo = p.use_tma ? egrp * (WS_TM / 2) : 0, m_hi = p.use_tma ? m_lo + WS_TM / 2 : WS_TM;
        const uint32_t tl = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        const bool has_res = (p.mode == SRM_TRUNK) | (p.mode == SRM_ADD_STORE_F16);
        unsigned it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
            const int y0 = p.y_lo + ty * SR_TY, x0 = tx * WS_TX;
            const unsigned a = it & 1;
            sr_mbar_wait(acc_full + a, (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (p.mode >= SRM_STORE_F16_SFT) {
                if (N >= 32) {
                    epilogue_sft<(N >= 32) ? N : 32>(p, sftw_s, sbias, tl + a * WS_TM * C::NACC, m_lo, m_hi, y0, x0, warp & 3, tid & 31);
                }
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                ws_arrive(acc_empty + a);
                continue;
            }
#pragma unroll 1
            for (int m = m_lo; m < m_hi; ++m) {
                const int py = y0 + (et >> 3), px = x0 + m * 8 + (et & 7);
                const bool inside = (py < p.y_hi) & (px < p.W);
                const int oy = py * p.out_s + p.out_oy, ox = px * p.out_s + p.out_ox;
                const size_t pix = (size_t)oy * ((size_t)p.W * p.out_s) + (size_t)ox;
                float4 res[NCH * 4];
                if (has_res && inside) {
                    const float4* ad = reinterpret_cast<const float4*>(p.add_f + pix * 64);
#pragma unroll
                    for (int q = 0; q < NCH * 4; ++q) res[q] = __ldg(ad + q);
                }
                uint32_t v[NCH][16];
#pragma unroll
                for (int c16 = 0; c16 < NCH; ++c16) sr_ld16(tl + (a * WS_TM + m) * C::NACC + c16 * 16, v[c16]);
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                if (!inside) continue;
#pragma unroll
                for (int c16 = 0; c16 < NCH; ++c16) {
                    float o[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[c16][j]) + sbias[c16 * 16 + j];
                    if (p.mode == SRM_STORE_F16) {
                        if (p.lrelu > 0.f) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) o[j] = o[j] > 0.f ? o[j] : o[j] * p.lrelu;
                        }
                        __half2 h[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) h[j] = sr_h2sat(o[2 * j], o[2 * j + 1]);
                        uint4* d = reinterpret_cast<uint4*>(p.dst_h + pix * p.dst_cstride + p.dst_c0 + c16 * 16);
                        d[0] = *reinterpret_cast<uint4*>(&h[0]);
                        d[1] = *reinterpret_cast<uint4*>(&h[4]);
                    } else if (p.mode == SRM_TRUNK) {
                        float4* d = reinterpret_cast<float4*>(p.dst_f + pix * 64 + c16 * 16);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 r = res[c16 * 4 + q];
                            d[q] = make_float4(o[4 * q] * p.scale + r.x, o[4 * q + 1] * p.scale + r.y, o[4 * q + 2] * p.scale + r.z, o[4 * q + 3] * p.scale + r.w);
                        }
                    } else if (p.mode == SRM_ADD_STORE_F16) {
                        __half2 h[8];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 r = res[c16 * 4 + q];
                            h[2 * q] = sr_h2sat(o[4 * q] + r.x, o[4 * q + 1] + r.y);
                            h[2 * q + 1] = sr_h2sat(o[4 * q + 2] + r.z, o[4 * q + 3] + r.w);
                        }
                        uint4* d = reinterpret_cast<uint4*>(p.dst_h + pix * p.dst_cstride + p.dst_c0 + c16 * 16);
                        d[0] = *reinterpret_cast<uint4*>(&h[0]);
                        d[1] = *reinterpret_cast<uint4*>(&h[4]);
                    } else if (p.mode == SRM_STORE_F32F16) {
                        float4* d = reinterpret_cast<float4*>(p.dst_f + pix * 64 + c16 * 16);
#pragma unroll
                        for (int q = 0; q < 4; ++q) d[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
                        if (p.dst_f2) {
                            float4* d2 = reinterpret_cast<float4*>(p.dst_f2 + pix * 64 + c16 * 16);
#pragma unroll
                            for (int q = 0; q < 4; ++q) d2[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
                        }
                    } else if (oy >= p.crop_y0 && oy < p.crop_y1 && ox >= p.crop_x0 && ox < p.crop_x1) {  // SRM_OUT_NCHW
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c16 * 16 + j < p.n_valid)
                                sr_store_out(p, (long long)(c16 * 16 + j) * p.out_ps + (long long)(oy - p.crop_y0) * p.out_rs + (ox - p.crop_x0), o[j]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            ws_arrive(acc_empty + a);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tbase), "r"((uint32_t)C::TCOLS) : "memory");
}

// ---------------------------------------------------------------------------------------------
// conv3x3_k64_kernel<N>: the 64-input-channel convolutions (conv_body, conv_up1/2 phases, conv_hr, conv_last --
// every layer after the dense blocks, i.e. all the high-resolution work) with 128-BYTE shared-memory rows.
// profiles/r1_conv_ws_ncu.md: those layers are bound by the copy engine's row rate -- conv3x3_ws_kernel moves a
// 16-channel slice as 612 rows of 16 B per plane.  Here a pixel's 64 channels are ONE 128-byte row in the
// SWIZZLE_128B K-major layout: the whole (16+2) x (24+2) halo of a tile is a single TMA tensor copy of 468 rows,
// the nine taps are still nothing but descriptor start addresses (+ (dy*26 + dx) * 128 B; tools/micro/
// conv_sw128_probe.cu shows that tap-shifted SW128 descriptors are exact with base_offset 0), a k16 step is
// +32 B inside the row, and -- because Cin = 64 is a single K slice -- the layer's weights ([tap][n][64] fp16,
// 74 KB for N = 64) are loaded ONCE per persistent CTA and stay resident.  Roles and barriers as in
// conv3x3_ws_kernel (producer thread / elected MMA lane / two epilogue warpgroups, two TMEM accumulator sets).
// ---------------------------------------------------------------------------------------------
constexpr int K64_TM = 3;                                       // 16 x 24 output pixels per tile
constexpr int K64_TX = 8 * K64_TM, K64_HX = K64_TX + 2;         // 24 / 26
constexpr int K64_ROWB = 128;                                   // bytes per pixel row (64 fp16)
constexpr int K64_A_BYTES = SR_HY * K64_HX * K64_ROWB;          // 59904: what one TMA copy delivers
constexpr int K64_A_STAGE = (K64_A_BYTES + 1023) / 1024 * 1024; // 60416
constexpr int K64_THREADS = 320;

template <int N> struct K64Cfg {
    static constexpr int NACC = N;                               // 64 or 16
    static constexpr int B_TAP = N * K64_ROWB;                   // 8192 / 2048 (1024-aligned)
    static constexpr int B_BYTES = 9 * B_TAP;
    static constexpr int NST = (N >= 64) ? 2 : 3;
    static constexpr int TCOLS = (2 * K64_TM * NACC <= 128) ? 128 : 512;   // 96 -> 128, 384 -> 512
    static constexpr int SMEM = 1024 + B_BYTES + NST * K64_A_STAGE + 512;   // alignment slack + B + ring + barriers/bias
};

__device__ __forceinline__ uint64_t k64_desc(uint32_t saddr, uint32_t sbo_bytes) {      // SWIZZLE_128B, K-major, base_offset 0
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <int N>
__global__ void __launch_bounds__(K64_THREADS, 1) conv3x3_k64_kernel(const __grid_constant__ ConvParams p, const __grid_constant__ CUtensorMap tmap_a,
                                                                      const __grid_constant__ CUtensorMap tmap_b) {
    using C = K64Cfg<N>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* Bw = smem;                                                    // [9][N][128 B] swizzled
    unsigned char* Ar = smem + C::B_BYTES;                                       // ring of NST halo stages
    uint64_t* full = reinterpret_cast<uint64_t*>(Ar + C::NST * K64_A_STAGE);    // [NST] count 1 + tx
    uint64_t* empty = full + C::NST;                                             // [NST] count 1 (tcgen05.commit)
    uint64_t* acc_full = empty + C::NST;                                         // [2]
    uint64_t* acc_empty = acc_full + 2;                                          // [2] count 256
    uint64_t* wbar = acc_empty + 2;                                              // weights landed
    uint32_t* tslot = reinterpret_cast<uint32_t*>(wbar + 1);
    float* sbias = reinterpret_cast<float*>(Ar + C::NST * K64_A_STAGE + 256);   // [64]
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid < 64) sbias[tid] = p.bias[tid];
    const int tiles_x = (p.W + K64_TX - 1) / K64_TX, tiles_y = (p.y_hi - p.y_lo + SR_TY - 1) / SR_TY;
    const int n_tiles = tiles_x * tiles_y;

    if (tid == 0) {
        for (int i = 0; i < C::NST; ++i) { sr_mbar_init(full + i, 1); sr_mbar_init(empty + i, 1); }
        sr_mbar_init(acc_full + 0, 1); sr_mbar_init(acc_full + 1, 1);
        sr_mbar_init(acc_empty + 0, 256); sr_mbar_init(acc_empty + 1, 256);
        sr_mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(sr_s32(tslot)), "r"((uint32_t)C::TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tbase = *tslot;

    if (warp == 9) {
        // ------------------------------ producer: weights once, then one tensor copy per tile ------------------------------
        if (tid == 288) {
            const uint64_t ta = reinterpret_cast<uint64_t>(&tmap_a), tb = reinterpret_cast<uint64_t>(&tmap_b);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(sr_s32(wbar)), "r"((uint32_t)p.ntaps * C::B_TAP) : "memory");
            for (int t = 0; t < p.ntaps; ++t)
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
                             :: "r"(sr_s32(Bw + t * C::B_TAP)), "l"(tb), "r"(sr_s32(wbar)), "r"(0), "r"(t * N) : "memory");
            SR_PDL_SYNC();                                        // the weights above are constants; the activations are not
            unsigned g = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++g) {
                const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
                const unsigned slot = g % C::NST;
                if (g >= (unsigned)C::NST) sr_mbar_wait(empty + slot, ((g / C::NST) - 1) & 1);
                const uint32_t bar = sr_s32(full + slot);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar), "r"((uint32_t)K64_A_BYTES) : "memory");
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                             :: "r"(sr_s32(Ar + slot * K64_A_STAGE)), "l"(ta), "r"(bar), "r"(0), "r"(tx * K64_TX - 1), "r"(p.y_lo + ty * SR_TY - 1) : "memory");
            }
        }
    } else if (warp == 8) {
        // ------------------------------ MMA issuer (one elected lane) ------------------------------
        if (sr_elect_one()) {
            const uint32_t idesc = sr_idesc(128, C::NACC);
            sr_mbar_wait(wbar, 0);
            unsigned g = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++g) {
                const unsigned a = g & 1, slot = g % C::NST;
                if (g >= 2) sr_mbar_wait(acc_empty + a, ((g >> 1) - 1) & 1);
                sr_mbar_wait(full + slot, (g / C::NST) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint64_t ad0 = k64_desc(sr_s32(Ar + slot * K64_A_STAGE), K64_HX * K64_ROWB);
                const uint64_t bd0 = k64_desc(sr_s32(Bw), 1024);
                const uint32_t dacc = tbase + a * K64_TM * C::NACC;
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    if (t < p.ntaps) {
                        const uint32_t aoff = (uint32_t)((p.tap_dy[t] * K64_HX + p.tap_dx[t]) * K64_ROWB) >> 4;
#pragma unroll
                        for (int m = 0; m < K64_TM; ++m)
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                sr_mma_ss(dacc + m * C::NACC, ad0 + (uint64_t)(aoff + m * 8 * (K64_ROWB >> 4) + ks * 2),
                                          bd0 + (uint64_t)(((t * C::B_TAP) >> 4) + ks * 2), idesc, (t | ks) != 0);
                    }
                }
                sr_commit(empty + slot);
                sr_commit(acc_full + a);
            }
        }
    } else {
        // ------------------------------ epilogue: 8 warps, thread = pixel, work items (m, 16-column block) dealt to the two groups ------------------------------
        SR_PDL_SYNC();                                            // residual reads / output writes touch earlier kernels' buffers
        constexpr int NCH = C::NACC / 16;
        constexpr int ITEMS = K64_TM * NCH;                     // 12 (N = 64) or 3 (N = 16)
        const int et = tid & 127, egrp = tid >> 7;
        const uint32_t tl = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        const bool has_res = (p.mode == SRM_ADD_STORE_F16);
        unsigned g = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++g) {
            const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
            const int y0 = p.y_lo + ty * SR_TY, x0 = tx * K64_TX;
            const unsigned a = g & 1;
            sr_mbar_wait(acc_full + a, (g >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll 1
            for (int it0 = egrp; it0 < ITEMS; it0 += 4) {        // two items of this group per round: it0 and it0 + 2
                uint32_t v[2][16];
                float4 res[2][4];
                bool ins[2]; size_t opix[2]; int c16s[2]; int oys[2], oxs[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int it = it0 + 2 * u;
                    const int m = (it < ITEMS) ? it / NCH : 0, c16 = (it < ITEMS) ? it % NCH : 0;
                    const int py = y0 + (et >> 3), px = x0 + m * 8 + (et & 7);
                    ins[u] = (it < ITEMS) & (py < p.y_hi) & (px < p.W);
                    oys[u] = py * p.out_s + p.out_oy; oxs[u] = px * p.out_s + p.out_ox;
                    opix[u] = (size_t)oys[u] * ((size_t)p.W * p.out_s) + (size_t)oxs[u];
                    c16s[u] = c16;
                    if (has_res && ins[u]) {
                        const float4* ad = reinterpret_cast<const float4*>(p.add_f + opix[u] * 64 + c16 * 16);
#pragma unroll
                        for (int q = 0; q < 4; ++q) res[u][q] = __ldg(ad + q);
                    }
                    if (it < ITEMS) sr_ld16(tl + (a * K64_TM + m) * C::NACC + c16 * 16, v[u]);     // warp-uniform condition
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (!ins[u]) continue;
                    const int c16 = c16s[u];
                    const size_t pix = opix[u];
                    float o[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[u][j]) + sbias[c16 * 16 + j];
                    if (p.mode == SRM_STORE_F16) {
                        if (p.lrelu > 0.f) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) o[j] = o[j] > 0.f ? o[j] : o[j] * p.lrelu;
                        }
                        __half2 h[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) h[j] = sr_h2sat(o[2 * j], o[2 * j + 1]);
                        uint4* d = reinterpret_cast<uint4*>(p.dst_h + pix * p.dst_cstride + p.dst_c0 + c16 * 16);
                        d[0] = *reinterpret_cast<uint4*>(&h[0]);
                        d[1] = *reinterpret_cast<uint4*>(&h[4]);
                    } else if (p.mode == SRM_ADD_STORE_F16) {
                        __half2 h[8];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 r = res[u][q];
                            h[2 * q] = sr_h2sat(o[4 * q] + r.x, o[4 * q + 1] + r.y);
                            h[2 * q + 1] = sr_h2sat(o[4 * q + 2] + r.z, o[4 * q + 3] + r.w);
                        }
                        uint4* d = reinterpret_cast<uint4*>(p.dst_h + pix * p.dst_cstride + p.dst_c0 + c16 * 16);
                        d[0] = *reinterpret_cast<uint4*>(&h[0]);
                        d[1] = *reinterpret_cast<uint4*>(&h[4]);
                    } else if (oys[u] >= p.crop_y0 && oys[u] < p.crop_y1 && oxs[u] >= p.crop_x0 && oxs[u] < p.crop_x1) {  // SRM_OUT_NCHW
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c16 * 16 + j < p.n_valid)
                                sr_store_out(p, (long long)(c16 * 16 + j) * p.out_ps + (long long)(oys[u] - p.crop_y0) * p.out_rs + (oxs[u] - p.crop_x0), o[j]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            ws_arrive(acc_empty + a);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tbase), "r"((uint32_t)C::TCOLS) : "memory");
}

// ---------------------------------------------------------------------------------------------
// SFT layer: scale = C1s(lrelu(C0s(cond))), shift = C1h(lrelu(C0h(cond))), y = x*(scale+1)+shift
// ---------------------------------------------------------------------------------------------
struct SftParams {
    const float* cond;               // [P,32] fp32
    const float* w;                  // packed: s0 [32][32], s0b[32], h0 [32][32], h0b[32], s1 [COUT][32], s1b, h1 [COUT][32], h1b
    const float* x_f; const __half* x_h; int xh_cstride, xh_c0;     // input: fp32 [P,64] or fp16 channel range
    __half* dst_h; int dst_cstride, dst_c0;                          // fp16 output (or nullptr)
    float* dst_f; const float* res_f; float res_scale;               // fp32 output: y*res_scale + res_f (RRDB tail) when dst_f != nullptr
    long long P;                                                     // pixels processed ...
    long long p0;                                                    // ... starting at this pixel (row window of a decoder unit)
};

template <int COUT>
__global__ void __launch_bounds__(128) sft_kernel(const __grid_constant__ SftParams p) {
    constexpr int NW = 2 * (32 * 32 + 32) + 2 * (COUT * 32 + COUT);
    extern __shared__ __align__(16) float sw[];
    for (int i = threadIdx.x; i < NW; i += blockDim.x) sw[i] = __ldg(p.w + i);
    __syncthreads();
    const float* s0 = sw;             const float* s0b = s0 + 1024;
    const float* h0 = s0b + 32;       const float* h0b = h0 + 1024;
    const float* s1 = h0b + 32;       const float* s1b = s1 + COUT * 32;
    const float* h1 = s1b + COUT;     const float* h1b = h1 + COUT * 32;
    const long long lp = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (lp >= p.P) return;
    const long long pix = p.p0 + lp;
    float c[32], ts[32], th[32];
    const float4* cp = reinterpret_cast<const float4*>(p.cond + pix * 32);
#pragma unroll
    for (int q = 0; q < 8; ++q) { const float4 v = __ldg(cp + q); c[4 * q] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w; }
#pragma unroll 4
    for (int o = 0; o < 32; ++o) {
        float a = s0b[o], b = h0b[o];
#pragma unroll
        for (int k = 0; k < 32; ++k) { a = fmaf(s0[o * 32 + k], c[k], a); b = fmaf(h0[o * 32 + k], c[k], b); }
        ts[o] = a > 0.f ? a : 0.2f * a;
        th[o] = b > 0.f ? b : 0.2f * b;
    }
#pragma unroll 1
    for (int o0 = 0; o0 < COUT; o0 += 8) {
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int o = o0 + j;
            float sc = s1b[o], sh = h1b[o];
#pragma unroll
            for (int k = 0; k < 32; ++k) { sc = fmaf(s1[o * 32 + k], ts[k], sc); sh = fmaf(h1[o * 32 + k], th[k], sh); }
            float x;
            if (p.x_f) x = __ldg(p.x_f + pix * 64 + o);
            else x = __half2float(p.x_h[pix * p.xh_cstride + p.xh_c0 + o]);
            y[j] = x * (sc + 1.f) + sh;
        }
        if (p.dst_f) {
#pragma unroll
            for (int j = 0; j < 8; ++j) p.dst_f[pix * 64 + o0 + j] = y[j] * p.res_scale + __ldg(p.res_f + pix * 64 + o0 + j);
        } else {
            __half2 h[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = sr_h2sat(y[2 * j], y[2 * j + 1]);
            *reinterpret_cast<uint4*>(p.dst_h + pix * p.dst_cstride + p.dst_c0 + o0) = *reinterpret_cast<uint4*>(h);
        }
    }
}


// ---------------------------------------------------------------------------------------------
// SFT layer on the tensor cores: both 1x1-conv branches are two chained K=32 GEMMs per 128-pixel
// tile -- L0: [cond] x [scale_conv0 ; shift_conv0] (N = 64), LeakyReLU + fp16 repack in TMEM,
// L1: scale = H[:, :32] x scale_conv1, shift = H[:, 32:] x shift_conv1 (A straight from TMEM) --
// and the modulation y = x*(scale+1)+shift is the epilogue (the +1 lives in the bias tile).
// Persistent CTAs: the layer's 16-22 KB operand blob is staged once per CTA.
// ---------------------------------------------------------------------------------------------
struct SftBlob { int off_b0, off_b1s, off_b1h, off_bias0, off_bias1s, off_bias1h, off_ones, total; };
__host__ __device__ inline SftBlob sft_blob_layout(int cout) {
    SftBlob L; int o = 0;
    L.off_b0 = o; o += 64 * 64;
    L.off_b1s = o; o += cout * 64;
    L.off_b1h = o; o += cout * 64;
    L.off_bias0 = o; o += 64 * 32;
    L.off_bias1s = o; o += cout * 32;
    L.off_bias1h = o; o += cout * 32;
    L.off_ones = o; o += 128 * 32;
    L.total = o;
    return L;
}

__device__ __forceinline__ void sr_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n}\n"
                 :: "r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void sr_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void sr_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}

struct SftTcParams {
    const float* cond;               // [P,32] fp32
    const unsigned char* blob;       // sft_blob_layout(COUT)
    const float* x_f; const __half* x_h; int xh_cstride, xh_c0;
    __half* dst_h; int dst_cstride, dst_c0;
    float* dst_f; const float* res_f; float res_scale;
    long long P; long long p0; int n_tiles;
};

template <int COUT>
__global__ void __launch_bounds__(128, 4) sft_tc_kernel(const __grid_constant__ SftTcParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const SftBlob BL = sft_blob_layout(COUT);
    unsigned char* blob = smem;
    unsigned char* atile = smem + ((BL.total + 1023) & ~1023);          // [128][32] fp16 canonical, 8 KB
    uint64_t* mbar = reinterpret_cast<uint64_t*>(atile + 8192);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(atile + 8192 + 16);
    // Tensor memory: 128 columns for either width, so four CTAs share an SM (the kernel is a chain of dependent phases
    // with 4 warps per CTA: occupancy is what hides its latencies).  The 64-channel layer runs its second GEMM in two
    // 32-channel halves: [0,32) packed fp16 hidden layer, [32,64) scale half, [64,96) shift half.
    constexpr int TCOLS = 128;
    constexpr int NH = 32;                                  // output channels per second-GEMM pass
    constexpr uint32_t D0 = 0, D1S = (COUT == 64) ? 32 : 64, D1H = (COUT == 64) ? 64 : 96;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < BL.total / 16; i += 128) cp_async16(blob + i * 16, p.blob + i * 16, 16);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    if (tid == 0) {
        sr_mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(sr_s32(tslot)), "r"((uint32_t)TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tbase = *tslot;
    const uint32_t tl = tbase + ((uint32_t)(warp * 32) << 16);
    const uint32_t blob_s = sr_s32(blob), a_s = sr_s32(atile);
    uint32_t ph = 0;
    SR_PDL_SYNC();

    // Register software pipeline: the kernel is a chain of dependent phases per 128-pixel tile (cond -> GEMM 1 ->
    // repack -> GEMM 2 -> modulate) with only 8-16 warps per SM, so every global load that is issued where it is
    // needed costs a full memory latency.  The condition row of tile i+1 is loaded while tile i is processed and
    // the x row of tile i is requested before its GEMMs, so both land behind the tensor-core phases.
    float4 cnd[8];
    auto load_cond = [&](int t) {
        const long long px_ = (long long)t * 128 + tid;
        const float4* cp = reinterpret_cast<const float4*>(p.cond + (p.p0 + (px_ < p.P ? px_ : 0)) * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) cnd[k] = __ldg(cp + k);
    };
    if ((int)blockIdx.x < p.n_tiles) load_cond(blockIdx.x);
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const bool valid = (long long)tile * 128 + tid < p.P;
        const long long pix = p.p0 + (long long)tile * 128 + tid;
        // cond row -> fp16, canonical layout (row = tid, 4 chunks of 8 channels)
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
            const float4 a = cnd[2 * kc], b = cnd[2 * kc + 1];
            __half2 h[4] = {sr_h2sat(a.x, a.y), sr_h2sat(a.z, a.w), sr_h2sat(b.x, b.y), sr_h2sat(b.z, b.w)};
            *reinterpret_cast<uint4*>(atile + tc_canon_off(tid, kc, 4)) = *reinterpret_cast<uint4*>(h);
        }
        if (tile + (int)gridDim.x < p.n_tiles) load_cond(tile + gridDim.x);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        if (warp == 0 && sr_elect_one()) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t id64 = sr_idesc(128, 64);
            sr_mma_ss(tbase + D0, sr_desc(a_s, 128, 512), sr_desc(blob_s + BL.off_b0, 128, 512), id64, 0);
            sr_mma_ss(tbase + D0, sr_desc(a_s + 256, 128, 512), sr_desc(blob_s + BL.off_b0 + 256, 128, 512), id64, 1);
            sr_mma_ss(tbase + D0, sr_desc(blob_s + BL.off_ones, 128, 256), sr_desc(blob_s + BL.off_bias0, 128, 256), id64, 1);
            sr_commit(mbar);
        }
        sr_mbar_wait(mbar, ph); ph ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
        for (int c = 0; c < 2; ++c) {   // LeakyReLU(0.2) + fp16 repack: H_scale -> cols [0,16), H_shift -> cols [16,32)
            uint32_t v[32], h[16];
            sr_ld32(tl + D0 + c * 32, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float a = __uint_as_float(v[2 * j]), b = __uint_as_float(v[2 * j + 1]);
                const __half2 hh = sr_h2sat(fmaxf(a, 0.2f * a), fmaxf(b, 0.2f * b));
                h[j] = *reinterpret_cast<const uint32_t*>(&hh);
            }
            sr_st16(tl + D0 + c * 16, h);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
#pragma unroll
        for (int hf = 0; hf < COUT / NH; ++hf) {
        uint4 xraw[NH / 4];                                     // this half of the x row: NH fp32 (all of it) or NH fp16 (first half);
        if (valid) {                                            // requested before the GEMM so that it lands behind it
            if (p.x_f) {
                const uint4* xp = reinterpret_cast<const uint4*>(p.x_f + pix * 64 + hf * NH);
#pragma unroll
                for (int q = 0; q < NH / 4; ++q) xraw[q] = __ldg(xp + q);
            } else {
                const uint4* xp = reinterpret_cast<const uint4*>(p.x_h + pix * p.xh_cstride + p.xh_c0 + hf * NH);
#pragma unroll
                for (int q = 0; q < NH / 8; ++q) xraw[q] = xp[q];
            }
        }
        if (warp == 0 && sr_elect_one()) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t idc = sr_idesc(128, NH);
            const uint32_t wo = (uint32_t)hf * (NH / 8) * 512, bo = (uint32_t)hf * (NH / 8) * 256;   // row blocks of 8 outputs: 512 B (K=32), 256 B (K=16)
            sr_mma_ts(tbase + D1S, tbase + D0 + 0, sr_desc(blob_s + BL.off_b1s + wo, 128, 512), idc, 0);
            sr_mma_ts(tbase + D1S, tbase + D0 + 8, sr_desc(blob_s + BL.off_b1s + wo + 256, 128, 512), idc, 1);
            sr_mma_ss(tbase + D1S, sr_desc(blob_s + BL.off_ones, 128, 256), sr_desc(blob_s + BL.off_bias1s + bo, 128, 256), idc, 1);
            sr_mma_ts(tbase + D1H, tbase + D0 + 16, sr_desc(blob_s + BL.off_b1h + wo, 128, 512), idc, 0);
            sr_mma_ts(tbase + D1H, tbase + D0 + 24, sr_desc(blob_s + BL.off_b1h + wo + 256, 128, 512), idc, 1);
            sr_mma_ss(tbase + D1H, sr_desc(blob_s + BL.off_ones, 128, 256), sr_desc(blob_s + BL.off_bias1h + bo, 128, 256), idc, 1);
            sr_commit(mbar);
        }
        sr_mbar_wait(mbar, ph); ph ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
        for (int cc = 0; cc < NH / 16; ++cc) {
            const int c16 = hf * (NH / 16) + cc;
            uint32_t sv[16], hv[16];
            sr_ld16(tl + D1S + cc * 16, sv);
            sr_ld16(tl + D1H + cc * 16, hv);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            if (!valid) continue;
            float x[16];
            if (p.x_f) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 t = xraw[cc * 4 + q];
                    x[4 * q] = __uint_as_float(t.x); x[4 * q + 1] = __uint_as_float(t.y); x[4 * q + 2] = __uint_as_float(t.z); x[4 * q + 3] = __uint_as_float(t.w);
                }
            } else {
                const uint4 u0 = xraw[cc * 2], u1 = xraw[cc * 2 + 1];
                const __half2* hp0 = reinterpret_cast<const __half2*>(&u0);
                const __half2* hp1 = reinterpret_cast<const __half2*>(&u1);
#pragma unroll
                for (int q = 0; q < 4; ++q) { const float2 a = __half22float2(hp0[q]), b = __half22float2(hp1[q]); x[2 * q] = a.x; x[2 * q + 1] = a.y; x[8 + 2 * q] = b.x; x[8 + 2 * q + 1] = b.y; }
            }
            float y[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) y[j] = fmaf(x[j], __uint_as_float(sv[j]), __uint_as_float(hv[j]));   // scale tile already holds scale+1
            if (p.dst_f) {
                const float4* rp = reinterpret_cast<const float4*>(p.res_f + pix * 64 + c16 * 16);
                float4* dp = reinterpret_cast<float4*>(p.dst_f + pix * 64 + c16 * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 r = __ldg(rp + q);
                    dp[q] = make_float4(y[4 * q] * p.res_scale + r.x, y[4 * q + 1] * p.res_scale + r.y, y[4 * q + 2] * p.res_scale + r.z, y[4 * q + 3] * p.res_scale + r.w);
                }
            } else {
                __half2 h[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) h[j] = sr_h2sat(y[2 * j], y[2 * j + 1]);
                uint4* d = reinterpret_cast<uint4*>(p.dst_h + pix * p.dst_cstride + p.dst_c0 + c16 * 16);
                d[0] = *reinterpret_cast<uint4*>(&h[0]);
                d[1] = *reinterpret_cast<uint4*>(&h[4]);
            }
        }
        if (hf + 1 < COUT / NH) {          // the next half overwrites D1S / D1H
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncthreads();
        }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();       // D1S/D1H and the A tile are free again
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tbase), "r"((uint32_t)TCOLS) : "memory");
}

// fp32 SFT weights (sft_kernel packing order s0,s0b,h0,h0b,s1,s1b,h1,h1b) -> tcgen05 operand blob
__global__ void pack_sft_blob_kernel(const float* __restrict__ w, unsigned char* __restrict__ blob, int cout) {
    const SftBlob BL = sft_blob_layout(cout);
    const float* s0 = w; const float* s0b = s0 + 1024; const float* h0 = s0b + 32; const float* h0b = h0 + 1024;
    const float* s1 = h0b + 32; const float* s1b = s1 + cout * 32; const float* h1 = s1b + cout; const float* h1b = h1 + cout * 32;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    auto put = [&](int off, int row, int k, int kch, float v) {
        *reinterpret_cast<__half*>(blob + off + tc_canon_off(row, k >> 3, kch) + (k & 7) * 2) = __float2half_rn(v);
    };
    auto put_bias = [&](int off, int row, float b) {
        const __half hi = __float2half_rn(b);
        put(off, row, 0, 2, __half2float(hi));
        put(off, row, 1, 2, b - __half2float(hi));
    };
    if (i < 64 * 32) { const int n = i >> 5, k = i & 31; put(BL.off_b0, n, k, 4, n < 32 ? s0[n * 32 + k] : h0[(n - 32) * 32 + k]); }
    if (i < cout * 32) { const int n = i >> 5, k = i & 31; put(BL.off_b1s, n, k, 4, s1[n * 32 + k]); put(BL.off_b1h, n, k, 4, h1[n * 32 + k]); }
    if (i < 64) put_bias(BL.off_bias0, i, i < 32 ? s0b[i] : h0b[i - 32]);
    if (i < cout) { put_bias(BL.off_bias1s, i, s1b[i] + 1.f); put_bias(BL.off_bias1h, i, h1b[i]); }
    if (i < 128) { put(BL.off_ones, i, 0, 2, 1.f); put(BL.off_ones, i, 1, 2, 1.f); }
}

// fp32 SFT weights (sft_kernel packing order) -> fragment-ordered block of the fused epilogues (sft_frag_layout)
__global__ void pack_sft_frag_kernel(const float* __restrict__ w, unsigned char* __restrict__ dst, int cout) {
    const SftFrag L = sft_frag_layout(cout);
    const float* s0 = w; const float* s0b = s0 + 1024; const float* h0 = s0b + 32; const float* h0b = h0 + 1024;
    const float* s1 = h0b + 32; const float* s1b = s1 + cout * 32; const float* h1 = s1b + cout; const float* h1b = h1 + cout * 32;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 64 * 32) {
        const int n = i >> 5, k = i & 31;
        reinterpret_cast<__half*>(dst + L.off_w0)[n * SFTF_PITCH + k] = __float2half_rn(n < 32 ? s0[n * 32 + k] : h0[(n - 32) * 32 + k]);
    }
    if (i < cout * 32) {
        const int n = i >> 5, k = i & 31;
        reinterpret_cast<__half*>(dst + L.off_w1s)[n * SFTF_PITCH + k] = __float2half_rn(s1[n * 32 + k]);
        reinterpret_cast<__half*>(dst + L.off_w1h)[n * SFTF_PITCH + k] = __float2half_rn(h1[n * 32 + k]);
    }
    if (i < 64) reinterpret_cast<float*>(dst + L.off_b0)[i] = i < 32 ? s0b[i] : h0b[i - 32];
    if (i < cout) { reinterpret_cast<float*>(dst + L.off_b1s)[i] = s1b[i] + 1.f; reinterpret_cast<float*>(dst + L.off_b1h)[i] = h1b[i]; }
}

// CondNet: conv3x3(1->64) lrelu, 1x1 64->64 lrelu, 1x1 64->64 lrelu, 1x1 64->32   (lib/sr_esrnet.py:440-444)
struct CondParams { const float* cond_in; const float* w; float* cond_out; int H, W; int y_lo, y_hi; __half* cond16; };   // w: c0 [64][9], b0[64], c2 [64][64], b2, c4 [64][64], b4, c6 [32][64], b6

__global__ void __launch_bounds__(128) condnet_kernel(const __grid_constant__ CondParams p) {
    constexpr int NW = 64 * 9 + 64 + 2 * (64 * 64 + 64) + 32 * 64 + 32;
    extern __shared__ __align__(16) float sw[];
    for (int i = threadIdx.x; i < NW; i += blockDim.x) sw[i] = __ldg(p.w + i);
    __syncthreads();
    const float* c0 = sw; const float* b0 = c0 + 576; const float* c2 = b0 + 64; const float* b2 = c2 + 4096;
    const float* c4 = b2 + 64; const float* b4 = c4 + 4096; const float* c6 = b4 + 64; const float* b6 = c6 + 2048;
    SR_PDL_SYNC();
    const long long pix = (long long)p.y_lo * p.W + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (long long)p.y_hi * p.W) return;
    const int y = (int)(pix / p.W), x = (int)(pix - (long long)y * p.W);
    float n[9];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int yy = y + dy - 1, xx = x + dx - 1;
            n[dy * 3 + dx] = (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) ? __ldg(p.cond_in + (size_t)yy * p.W + xx) : 0.f;
        }
    float a[64], b[64];
#pragma unroll 4
    for (int o = 0; o < 64; ++o) {
        float v = b0[o];
#pragma unroll
        for (int k = 0; k < 9; ++k) v = fmaf(c0[o * 9 + k], n[k], v);
        a[o] = v > 0.f ? v : 0.2f * v;
    }
#pragma unroll 2
    for (int o = 0; o < 64; ++o) {
        float v = b2[o];
#pragma unroll
        for (int k = 0; k < 64; ++k) v = fmaf(c2[o * 64 + k], a[k], v);
        b[o] = v > 0.f ? v : 0.2f * v;
    }
#pragma unroll 2
    for (int o = 0; o < 64; ++o) {
        float v = b4[o];
#pragma unroll
        for (int k = 0; k < 64; ++k) v = fmaf(c4[o * 64 + k], b[k], v);
        a[o] = v > 0.f ? v : 0.2f * v;
    }
#pragma unroll 2
    for (int o = 0; o < 32; ++o) {
        float v = b6[o];
#pragma unroll
        for (int k = 0; k < 64; ++k) v = fmaf(c6[o * 64 + k], a[k], v);
        p.cond_out[pix * 32 + o] = v;
        if (p.cond16) p.cond16[pix * 32 + o] = __float2half_rn(v);         // operand of the fused SFT epilogues
    }
}

// ---------------------------------------------------------------------------------------------
// CondNet on the tensor cores (round 2): the 1x1 layers (64->64, 64->64, 64->32) as mma.sync GEMMs on 16-pixel row
// blocks, fp32-class through a 3-term fp16 split (a = ah + al, W = Wh + Wl; ah*Wh + al*Wh + ah*Wl, fp32 accumulate:
// the dropped al*Wl term is ~2^-22 relative), the 3x3 layer (9 MACs per output) in fp32 FFMA computed directly in the
// A-fragment layout.  The condition map feeds every SFT layer, so it is kept at fp32 accuracy; the fp32 FFMA kernel
// above (254 us per 520x520 tile, 4 % of the decoder) stays as the reference (K4_CONDNET_FP32=1).
// ---------------------------------------------------------------------------------------------
constexpr int CNM_PITCH = 72;                                   // halfs per weight row (64 + 8: conflict-free fragment loads)
struct CnmLayout { int w1h, w1l, w2h, w2l, w3h, w3l, w0, b0, b1, b2, b3, total; };
__host__ __device__ inline CnmLayout cnm_layout() {
    CnmLayout L; int o = 0;
    L.w1h = o; o += 64 * CNM_PITCH * 2; L.w1l = o; o += 64 * CNM_PITCH * 2;
    L.w2h = o; o += 64 * CNM_PITCH * 2; L.w2l = o; o += 64 * CNM_PITCH * 2;
    L.w3h = o; o += 32 * CNM_PITCH * 2; L.w3l = o; o += 32 * CNM_PITCH * 2;
    L.w0 = o; o += 64 * 9 * 4; L.b0 = o; o += 64 * 4; L.b1 = o; o += 64 * 4; L.b2 = o; o += 64 * 4; L.b3 = o; o += 32 * 4;
    L.total = (o + 15) & ~15;
    return L;
}
// w: c0 [64][9], b0[64], c2 [64][64], b2, c4 [64][64], b4, c6 [32][64], b6  (the fp32 kernel's packing)
__global__ void pack_condnet_mma_kernel(const float* __restrict__ w, unsigned char* __restrict__ dst) {
    const CnmLayout L = cnm_layout();
    const float* c0 = w; const float* b0 = c0 + 576; const float* c2 = b0 + 64; const float* b2 = c2 + 4096;
    const float* c4 = b2 + 64; const float* b4 = c4 + 4096; const float* c6 = b4 + 64; const float* b6 = c6 + 2048;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    auto split = [&](int offh, int offl, int n, int k, float v) {
        const __half hi = __float2half_rn(v);
        reinterpret_cast<__half*>(dst + offh)[n * CNM_PITCH + k] = hi;
        reinterpret_cast<__half*>(dst + offl)[n * CNM_PITCH + k] = __float2half_rn(v - __half2float(hi));
    };
    if (i < 4096) { split(L.w1h, L.w1l, i >> 6, i & 63, c2[i]); split(L.w2h, L.w2l, i >> 6, i & 63, c4[i]); }
    if (i < 2048) split(L.w3h, L.w3l, i >> 6, i & 63, c6[i]);
    if (i < 576) reinterpret_cast<float*>(dst + L.w0)[i] = c0[i];
    if (i < 64) { reinterpret_cast<float*>(dst + L.b0)[i] = b0[i]; reinterpret_cast<float*>(dst + L.b1)[i] = b2[i]; reinterpret_cast<float*>(dst + L.b2)[i] = b4[i]; }
    if (i < 32) reinterpret_cast<float*>(dst + L.b3)[i] = b6[i];
}

struct CondMmaParams { const float* cond_in; const unsigned char* wblk; float* cond_out; __half* cond16; int H, W, y_lo, y_hi; };

// accumulator fragments of 2 adjacent n8 blocks -> hi / lo A fragments of one k16 step
__device__ __forceinline__ void cnm_split_frag(const float (&c0)[4], const float (&c1)[4], uint32_t (&ah)[4], uint32_t (&al)[4]) {
    auto pk = [](float x, float y, uint32_t& hi, uint32_t& lo) {
        const __half2 h = __floats2half2_rn(x, y);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
        hi = *reinterpret_cast<const uint32_t*>(&h); lo = *reinterpret_cast<const uint32_t*>(&l);
    };
    pk(c0[0], c0[1], ah[0], al[0]); pk(c0[2], c0[3], ah[1], al[1]);
    pk(c1[0], c1[1], ah[2], al[2]); pk(c1[2], c1[3], ah[3], al[3]);
}
// one 64-input layer for 16 pixels: NBLK n8 blocks of outputs, 3-term split products
template <int NBLK>
__device__ __forceinline__ void cnm_layer(const unsigned char* sm, int offh, int offl, int offb, const uint32_t (&ah)[4][4], const uint32_t (&al)[4][4],
                                          int g, int tq, float (&c)[NBLK][4]) {
    const __half* Wh = reinterpret_cast<const __half*>(sm + offh);
    const __half* Wl = reinterpret_cast<const __half*>(sm + offl);
    const float* b = reinterpret_cast<const float*>(sm + offb);
#pragma unroll
    for (int j = 0; j < NBLK; ++j) {
        c[j][0] = c[j][2] = b[8 * j + 2 * tq]; c[j][1] = c[j][3] = b[8 * j + 2 * tq + 1];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const __half* wh = Wh + (8 * j + g) * CNM_PITCH + 16 * ks + 2 * tq;
            const __half* wl = Wl + (8 * j + g) * CNM_PITCH + 16 * ks + 2 * tq;
            const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(wh), bh1 = *reinterpret_cast<const uint32_t*>(wh + 8);
            sr_hmma(c[j], al[ks], bh0, bh1);                     // small terms first
            sr_hmma(c[j], ah[ks], *reinterpret_cast<const uint32_t*>(wl), *reinterpret_cast<const uint32_t*>(wl + 8));
            sr_hmma(c[j], ah[ks], bh0, bh1);
        }
    }
}

__global__ void __launch_bounds__(128) condnet_mma_kernel(const __grid_constant__ CondMmaParams p) {
    extern __shared__ __align__(16) unsigned char csm[];
    const CnmLayout L = cnm_layout();
    for (int i = threadIdx.x; i < L.total / 16; i += blockDim.x) reinterpret_cast<uint4*>(csm)[i] = __ldg(reinterpret_cast<const uint4*>(p.wblk) + i);
    __syncthreads();
    SR_PDL_SYNC();
    const int lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const long long p0 = (long long)p.y_lo * p.W, pend = (long long)p.y_hi * p.W;
    const long long nblk = (pend - p0 + 15) / 16;
    const float* W0 = reinterpret_cast<const float*>(csm + L.w0);
    const float* B0 = reinterpret_cast<const float*>(csm + L.b0);
    for (long long blk = (long long)blockIdx.x * 4 + (threadIdx.x >> 5); blk < nblk; blk += (long long)gridDim.x * 4) {
        const long long pixA = p0 + blk * 16 + g, pixB = pixA + 8;
        const bool inA = pixA < pend, inB = pixB < pend;
        // 3x3 neighbourhoods of the two pixels (zero padding at the tile border, as the reference's conv)
        float nA[9], nB[9];
        {
            const long long pa = inA ? pixA : p0, pb = inB ? pixB : p0;
            const int ya = (int)(pa / p.W), xa = (int)(pa - (long long)ya * p.W), yb = (int)(pb / p.W), xb = (int)(pb - (long long)yb * p.W);
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int y1 = ya + dy - 1, x1 = xa + dx - 1, y2 = yb + dy - 1, x2 = xb + dx - 1;
                    nA[dy * 3 + dx] = (y1 >= 0 && y1 < p.H && x1 >= 0 && x1 < p.W) ? __ldg(p.cond_in + (size_t)y1 * p.W + x1) : 0.f;
                    nB[dy * 3 + dx] = (y2 >= 0 && y2 < p.H && x2 >= 0 && x2 < p.W) ? __ldg(p.cond_in + (size_t)y2 * p.W + x2) : 0.f;
                }
        }
        // layer 0 (3x3, 1 -> 64) in fp32, each thread the 16 channels of its A-fragment slots: k-step ks holds channels
        // 16ks + 2tq, +1 (regs 0 / 1: pixel A / B) and 16ks + 8 + 2tq, +1 (regs 2 / 3)
        uint32_t ah[4][4], al[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            float va[4], vb[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int ch = 16 * ks + (e >> 1) * 8 + 2 * tq + (e & 1);
                float a = B0[ch], b = B0[ch];
#pragma unroll
                for (int k = 0; k < 9; ++k) { a = fmaf(W0[ch * 9 + k], nA[k], a); b = fmaf(W0[ch * 9 + k], nB[k], b); }
                va[e] = fmaxf(a, 0.2f * a); vb[e] = fmaxf(b, 0.2f * b);
            }
            const float c0[4] = {va[0], va[1], vb[0], vb[1]}, c1[4] = {va[2], va[3], vb[2], vb[3]};
            cnm_split_frag(c0, c1, ah[ks], al[ks]);
        }
        float c[8][4];
        cnm_layer<8>(csm, L.w1h, L.w1l, L.b1, ah, al, g, tq, c);
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) c[j][e] = fmaxf(c[j][e], 0.2f * c[j][e]);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) cnm_split_frag(c[2 * ks], c[2 * ks + 1], ah[ks], al[ks]);
        cnm_layer<8>(csm, L.w2h, L.w2l, L.b2, ah, al, g, tq, c);
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) c[j][e] = fmaxf(c[j][e], 0.2f * c[j][e]);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) cnm_split_frag(c[2 * ks], c[2 * ks + 1], ah[ks], al[ks]);
        float o[4][4];
        cnm_layer<4>(csm, L.w3h, L.w3l, L.b3, ah, al, g, tq, o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ch = 8 * j + 2 * tq;
            if (inA) {
                *reinterpret_cast<float2*>(p.cond_out + pixA * 32 + ch) = make_float2(o[j][0], o[j][1]);
                if (p.cond16) *reinterpret_cast<__half2*>(p.cond16 + pixA * 32 + ch) = __floats2half2_rn(o[j][0], o[j][1]);
            }
            if (inB) {
                *reinterpret_cast<float2*>(p.cond_out + pixB * 32 + ch) = make_float2(o[j][2], o[j][3]);
                if (p.cond16) *reinterpret_cast<__half2*>(p.cond16 + pixB * 32 + ch) = __floats2half2_rn(o[j][2], o[j][3]);
            }
        }
    }
}

// planar fp32 [C,H,W] -> NHWC fp16 [H,W,32] (zero padded)
__global__ void nchw_to_nhwc32_kernel(const float* __restrict__ src, __half* __restrict__ dst, int C, long long P) {
    SR_PDL_SYNC();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * 32) return;
    const long long pix = i >> 5;
    const int c = (int)(i & 31);
    dst[i] = __float2half_rn(c < C ? src[(size_t)c * P + pix] : 0.f);
}

// conv weight [Cout][Cin][3][3] fp32 -> [cin_slice][tap][NPAD][SR_CK] canonical K-major fp16 tiles
__global__ void pack_conv3x3_kernel(const float* __restrict__ W, unsigned char* __restrict__ dst,
                                    int cout, int cin, int npad, int nslices) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)nslices * 9 * npad * SR_CK;
    if (i >= total) return;
    const int k = (int)(i % SR_CK);
    long long r = i / SR_CK;
    const int n = (int)(r % npad); r /= npad;
    const int t = (int)(r % 9);
    const int sl = (int)(r / 9);
    const int ci = sl * SR_CK + k;
    const float w = (n < cout && ci < cin) ? W[((size_t)n * cin + ci) * 9 + t] : 0.f;
    const size_t off = ((size_t)sl * 9 + t) * (npad * SR_CK * 2) + tc_canon_off(n, k >> 3, SR_CK / 8) + (k & 7) * 2;
    *reinterpret_cast<__half*>(dst + off) = __float2half_rn(w);
}

// Sub-pixel weights of "nearest x2 upsample, then 3x3 conv": output pixel (2Y+py, 2X+px) only sees the 2x2
// source pixels (Y+py-1..Y+py, X+px-1..X+px); each of them collects the kernel rows / columns that land on it
// after upsampling: phase 0: {-1: k0, 0: k1+k2}, phase 1: {0: k0+k1, +1: k2}.  Packed as [slice][4 taps][NPAD][SR_CK]
// (tap = 2*iy + ix); the sums are formed in fp32 and rounded to fp16 once.  4/9 of the MACs of the direct form.
__global__ void pack_conv_phase_kernel(const float* __restrict__ W, unsigned char* __restrict__ dst,
                                       int cout, int cin, int npad, int nslices, int py, int px) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)nslices * 4 * npad * SR_CK;
    if (i >= total) return;
    const int k = (int)(i % SR_CK);
    long long r = i / SR_CK;
    const int n = (int)(r % npad); r /= npad;
    const int t = (int)(r % 4);
    const int sl = (int)(r / 4);
    const int ci = sl * SR_CK + k;
    const int iy = t >> 1, ix = t & 1;
    // kernel index range collected by source offset (phase + i - 1)
    const int ky0 = (py == 0) ? (iy == 0 ? 0 : 1) : (iy == 0 ? 0 : 2), ky1 = (py == 0) ? (iy == 0 ? 0 : 2) : (iy == 0 ? 1 : 2);
    const int kx0 = (px == 0) ? (ix == 0 ? 0 : 1) : (ix == 0 ? 0 : 2), kx1 = (px == 0) ? (ix == 0 ? 0 : 2) : (ix == 0 ? 1 : 2);
    float w = 0.f;
    if (n < cout && ci < cin)
        for (int ky = ky0; ky <= ky1; ++ky)
            for (int kx = kx0; kx <= kx1; ++kx) w += W[((size_t)n * cin + ci) * 9 + ky * 3 + kx];
    const size_t off = ((size_t)sl * 4 + t) * (npad * SR_CK * 2) + tc_canon_off(n, k >> 3, SR_CK / 8) + (k & 7) * 2;
    *reinterpret_cast<__half*>(dst + off) = __float2half_rn(w);
}

// Weights for conv3x3_k64_kernel: plain [tap][NPAD][64] fp16 rows (the TMA copy applies the 128-byte swizzle).
// phase < 0: the nine taps of the direct 3x3 conv; phase 0..3: the four taps of that sub-pixel phase
// (same combination rule as pack_conv_phase_kernel).
__global__ void pack_conv_k64_kernel(const float* __restrict__ W, __half* __restrict__ dst, int cout, int cin, int npad, int phase) {
    const int ntaps = (phase < 0) ? 9 : 4;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)ntaps * npad * 64) return;
    const int k = (int)(i % 64);
    const int n = (int)((i / 64) % npad);
    const int t = (int)(i / 64 / npad);
    float w = 0.f;
    if (n < cout && k < cin) {
        if (phase < 0) {
            w = W[((size_t)n * cin + k) * 9 + t];
        } else {
            const int py = phase >> 1, px = phase & 1, iy = t >> 1, ix = t & 1;
            const int ky0 = (py == 0) ? (iy == 0 ? 0 : 1) : (iy == 0 ? 0 : 2), ky1 = (py == 0) ? (iy == 0 ? 0 : 2) : (iy == 0 ? 1 : 2);
            const int kx0 = (px == 0) ? (ix == 0 ? 0 : 1) : (ix == 0 ? 0 : 2), kx1 = (px == 0) ? (ix == 0 ? 0 : 2) : (ix == 0 ? 1 : 2);
            for (int ky = ky0; ky <= ky1; ++ky)
                for (int kx = kx0; kx <= kx1; ++kx) w += W[((size_t)n * cin + k) * 9 + ky * 3 + kx];
        }
    }
    dst[i] = __float2half_rn(w);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct SrConv { unsigned char* wpack; float* bias; int cin_pad, cout, npad; unsigned char* wphase[4];
                __half* wk64; __half* wk64_phase[4]; };       // conv3x3_k64_kernel packs (Cin == 64 layers only)
struct SrSft { float* w; unsigned char* blob; unsigned char* frag; int cout; };

struct k4_srnet {
    int num_feat, num_block, num_grow, num_cond, n_in, scale;
    SrConv conv_first, conv_body, conv_up1, conv_up2, conv_hr, conv_last;
    SrConv rdb_conv[32][3][5];
    SrSft rdb_sft[32][3][2];
    SrSft rrdb_sft[32];
    SrSft sftbody;
    float* condnet_w;
    unsigned char* condnet_mma;                 // cnm_layout block of condnet_mma_kernel
    void* allocs[1024];
    int n_allocs;
    size_t bytes;
};

namespace {

int sr_alloc(k4_srnet* n, void** p, size_t bytes) {
    if (n->n_allocs >= 1024) return K4_ERR_INVALID_ARG;
    K4_CUDA_TRY(cudaMalloc(p, bytes ? bytes : 16));
    n->allocs[n->n_allocs++] = *p;
    n->bytes += bytes;
    return K4_OK;
}

int make_conv(k4_srnet* n, SrConv& c, const float* w, const float* b, int cout, int cin, cudaStream_t s) {
    c.cout = cout;
    c.npad = cout <= 16 ? 16 : (cout <= 32 ? 32 : 64);
    c.cin_pad = (cin + 31) / 32 * 32;              // activations are stored in multiples of 32 channels
    const int nsl = c.cin_pad / SR_CK;
    const size_t bytes = (size_t)nsl * 9 * c.npad * SR_CK * 2;
    int st = sr_alloc(n, (void**)&c.wpack, bytes);
    if (st) return st;
    st = sr_alloc(n, (void**)&c.bias, 64 * 4);
    if (st) return st;
    K4_CUDA_TRY(cudaMemsetAsync(c.bias, 0, 64 * 4, s));
    K4_CUDA_TRY(cudaMemcpyAsync(c.bias, b, (size_t)cout * 4, cudaMemcpyDeviceToDevice, s));
    const long long total = (long long)nsl * 9 * c.npad * SR_CK;
    pack_conv3x3_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(w, c.wpack, cout, cin, c.npad, nsl);
    K4_CUDA_TRY(cudaGetLastError());
    return K4_OK;
}

// the four sub-pixel weight packs of a conv that follows a nearest-x2 upsampling (conv_up1 / conv_up2)
int make_conv_phases(k4_srnet* n, SrConv& c, const float* w, int cout, int cin, cudaStream_t s) {
    const int nsl = c.cin_pad / SR_CK;
    const size_t bytes = (size_t)nsl * 4 * c.npad * SR_CK * 2;
    const long long total = (long long)nsl * 4 * c.npad * SR_CK;
    for (int ph = 0; ph < 4; ++ph) {
        int st = sr_alloc(n, (void**)&c.wphase[ph], bytes);
        if (st) return st;
        pack_conv_phase_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(w, c.wphase[ph], cout, cin, c.npad, nsl, ph >> 1, ph & 1);
        K4_CUDA_TRY(cudaGetLastError());
    }
    return K4_OK;
}

// [tap][npad][64] packs for conv3x3_k64_kernel (direct 9 taps, and the 4 phases when `phases`)
int make_conv_k64(k4_srnet* n, SrConv& c, const float* w, int cout, int cin, bool phases, cudaStream_t s) {
    if (cin != 64 || (c.npad != 64 && c.npad != 16)) return K4_OK;
    {
        const long long total = 9LL * c.npad * 64;
        int st = sr_alloc(n, (void**)&c.wk64, (size_t)total * 2);
        if (st) return st;
        pack_conv_k64_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(w, c.wk64, cout, cin, c.npad, -1);
        K4_CUDA_TRY(cudaGetLastError());
    }
    for (int ph = 0; phases && ph < 4; ++ph) {
        const long long total = 4LL * c.npad * 64;
        int st = sr_alloc(n, (void**)&c.wk64_phase[ph], (size_t)total * 2);
        if (st) return st;
        pack_conv_k64_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(w, c.wk64_phase[ph], cout, cin, c.npad, ph);
        K4_CUDA_TRY(cudaGetLastError());
    }
    return K4_OK;
}

// SFT weights: 4 convs in the order scale0, scale1, shift0, shift1 (weight, bias each) -> packed s0,s0b,h0,h0b,s1,s1b,h1,h1b
int make_sft(k4_srnet* n, SrSft& f, const float* const* pw, int cout, cudaStream_t s) {
    f.cout = cout;
    const size_t nw = 2 * (1024 + 32) + 2 * ((size_t)cout * 32 + cout);
    int st = sr_alloc(n, (void**)&f.w, nw * 4);
    if (st) return st;
    float* d = f.w;
    auto cp = [&](const float* src, size_t cnt) { cudaMemcpyAsync(d, src, cnt * 4, cudaMemcpyDeviceToDevice, s); d += cnt; };
    cp(pw[0], 1024); cp(pw[1], 32);                 // scale_conv0 w, b
    cp(pw[4], 1024); cp(pw[5], 32);                 // shift_conv0 w, b
    cp(pw[2], (size_t)cout * 32); cp(pw[3], cout);  // scale_conv1 w, b
    cp(pw[6], (size_t)cout * 32); cp(pw[7], cout);  // shift_conv1 w, b
    K4_CUDA_TRY(cudaGetLastError());
    const SftBlob BL = sft_blob_layout(cout);
    st = sr_alloc(n, (void**)&f.blob, (size_t)BL.total);
    if (st) return st;
    K4_CUDA_TRY(cudaMemsetAsync(f.blob, 0, BL.total, s));
    pack_sft_blob_kernel<<<8, 256, 0, s>>>(f.w, f.blob, cout);
    K4_CUDA_TRY(cudaGetLastError());
    const SftFrag FL = sft_frag_layout(cout);
    st = sr_alloc(n, (void**)&f.frag, (size_t)FL.total);
    if (st) return st;
    K4_CUDA_TRY(cudaMemsetAsync(f.frag, 0, FL.total, s));
    pack_sft_frag_kernel<<<8, 256, 0, s>>>(f.w, f.frag, cout);
    K4_CUDA_TRY(cudaGetLastError());
    return K4_OK;
}

// K4_SR_PDL=0 launches the decoder kernels without programmatic dependent launch (A/B switch)
static bool sr_use_pdl() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("K4_SR_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

template <typename... KArgs, typename... Args>
cudaError_t sr_launch(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = sr_use_pdl() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// per-device caches of the launchers (function attributes are per device)
struct SrDevInfo { int sms; unsigned attr_mask; };
static SrDevInfo* sr_dev() {
    static SrDevInfo info[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (info[dev].sms == 0) cudaDeviceGetAttribute(&info[dev].sms, cudaDevAttrMultiProcessorCount, dev);
    return &info[dev];
}
template <typename F>
int sr_set_smem(F kern, int bit, int bytes) {
    SrDevInfo* d = sr_dev();
    if (!(d->attr_mask & (1u << bit))) {
        K4_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        d->attr_mask |= (1u << bit);
    }
    return K4_OK;
}
constexpr int sr_bit_n(int n) { return n == 64 ? 0 : (n == 32 ? 1 : 2); }

template <int N>
int launch_conv(const ConvParams& p, cudaStream_t s) {
    constexpr int STAGE = SR_A_STAGE + 9 * N * SR_CK * 2;
    constexpr int SMEM = 2 * STAGE + 64;
    if (int st = sr_set_smem(conv3x3_tc_kernel<N>, 0 + sr_bit_n(N), SMEM)) return st;
    const int tiles_y = (p.y_hi - p.y_lo + SR_TY - 1) / SR_TY;
    K4_CUDA_TRY(sr_launch(conv3x3_tc_kernel<N>, (unsigned)(p.tiles_x * tiles_y), 128, SMEM, s, p));
    return K4_OK;
}

typedef CUresult (*k4_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
static k4_encode_tiled_fn sr_encode_tiled() {
    static k4_encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<k4_encode_tiled_fn>(ptr);
        else
            cudaGetLastError();
    }
    return fn;
}

// (channel, x, y) view of the NHWC fp16 source starting at channel src_c0; box = one 8-channel halo plane
static bool sr_make_tmap(const ConvParams& p, CUtensorMap* tm) {
    k4_encode_tiled_fn enc = sr_encode_tiled();
    if (!enc) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)p.cin, (cuuint64_t)p.W, (cuuint64_t)p.H};
    const cuuint64_t gstride[2] = {(cuuint64_t)p.src_cstride * 2, (cuuint64_t)p.W * p.src_cstride * 2};
    const cuuint32_t box[3] = {8, (cuuint32_t)WS_HX, (cuuint32_t)SR_HY};
    const cuuint32_t estr[3] = {1, 1, 1};
    void* base = const_cast<__half*>(p.src + p.src_c0);
    if (((uintptr_t)base & 15) || (gstride[0] & 15)) return false;
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool sr_no_tma() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("K4_CONV_NO_TMA"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

template <int N>
int launch_conv_ws(ConvParams p, cudaStream_t s) {
    using C = WsCfg<N>;
    if (int st = sr_set_smem(conv3x3_ws_kernel<N>, 3 + sr_bit_n(N), C::SMEM + SFTF_MAX)) return st;
    const int sms = sr_dev()->sms;
    alignas(64) CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    p.use_tma = (!p.upsample && !sr_no_tma() && sr_make_tmap(p, &tm)) ? 1 : 0;
    const int tiles = ((p.W + WS_TX - 1) / WS_TX) * ((p.y_hi - p.y_lo + SR_TY - 1) / SR_TY);
    if (tiles <= 0) return K4_OK;
    size_t smem = C::SMEM;
    if (p.mode >= SRM_STORE_F16_SFT) {
        if (!p.use_tma || N < 32 || !p.cond16 || !p.sftw) return K4_ERR_UNSUPPORTED;     // fused epilogues: TMA mode (8 epilogue warps) only
        smem += sft_frag_layout(p.sft_n).total + (p.mode == SRM_TRUNK_SFT2 ? sft_frag_layout(64).total : 0);
    }
    K4_CUDA_TRY(sr_launch(conv3x3_ws_kernel<N>, (unsigned)(tiles < sms ? tiles : sms), WS_THREADS, smem, s, p, tm));
    return K4_OK;
}

// conv3x3_k64_kernel launch: tensor maps for the activation (channel, x, y; box 64 x 26 x 18) and the weight rows
template <int N>
int launch_conv_k64(ConvParams p, const __half* wrows, cudaStream_t s) {
    using C = K64Cfg<N>;
    k4_encode_tiled_fn enc = sr_encode_tiled();
    if (!enc) return K4_ERR_UNSUPPORTED;
    if (int st = sr_set_smem(conv3x3_k64_kernel<N>, 6 + sr_bit_n(N), C::SMEM)) return st;
    const int sms = sr_dev()->sms;
    alignas(64) CUtensorMap ta, tb;
    {
        const cuuint64_t gdim[3] = {64, (cuuint64_t)p.W, (cuuint64_t)p.H};
        const cuuint64_t gstr[2] = {(cuuint64_t)p.src_cstride * 2, (cuuint64_t)p.W * p.src_cstride * 2};
        const cuuint32_t box[3] = {64, (cuuint32_t)K64_HX, (cuuint32_t)SR_HY}, es[3] = {1, 1, 1};
        if (enc(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(p.src + p.src_c0), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return K4_ERR_UNSUPPORTED;
    }
    {
        const cuuint64_t gdim[2] = {64, (cuuint64_t)p.ntaps * N};
        const cuuint64_t gstr[1] = {128};
        const cuuint32_t box[2] = {64, (cuuint32_t)N}, es[2] = {1, 1};
        if (enc(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(wrows), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return K4_ERR_UNSUPPORTED;
    }
    const int tiles = ((p.W + K64_TX - 1) / K64_TX) * ((p.y_hi - p.y_lo + SR_TY - 1) / SR_TY);
    if (tiles <= 0) return K4_OK;
    K4_CUDA_TRY(sr_launch(conv3x3_k64_kernel<N>, (unsigned)(tiles < sms ? tiles : sms), K64_THREADS, C::SMEM, s, p, ta, tb));
    return K4_OK;
}

// The 64-input-channel layers go to conv3x3_k64_kernel; K4_CONV_K64=0 keeps them on conv3x3_ws_kernel (A/B switch).
static bool sr_use_k64() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("K4_CONV_K64"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

// K4_CONV_V1=1 selects the one-tile-per-CTA kernel (kept as the A/B reference of the persistent one)
static bool sr_use_v1() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("K4_CONV_V1"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

int run_conv(const SrConv& c, ConvParams p, cudaStream_t s, const __half* wk64_override = nullptr) {
    const __half* wk64 = wk64_override ? wk64_override : ((p.ntaps == 0 || p.ntaps == 9) ? c.wk64 : nullptr);
    if (!p.wpack) p.wpack = c.wpack;
    p.bias = c.bias; p.cin = c.cin_pad;
    if (p.ntaps == 0) {
        p.ntaps = 9;
        for (int t = 0; t < 9; ++t) { p.tap_dy[t] = (unsigned char)(t / 3); p.tap_dx[t] = (unsigned char)(t % 3); }
    }
    if (p.out_s == 0) p.out_s = 1;
    if (p.y_hi == 0) { p.y_lo = 0; p.y_hi = p.H; }
    if (p.y_lo < 0) p.y_lo = 0;
    if (p.y_hi > p.H) p.y_hi = p.H;
    if (p.y_hi <= p.y_lo) return K4_OK;
    p.tiles_x = (p.W + SR_TX - 1) / SR_TX;
    if (sr_use_v1()) {
        if (c.npad == 64) return launch_conv<64>(p, s);
        if (c.npad == 32) return launch_conv<32>(p, s);
        return launch_conv<16>(p, s);
    }
    if (sr_use_k64() && wk64 && p.cin == 64 && !p.upsample && (p.src_cstride % 8) == 0 && (p.src_c0 % 8) == 0 &&
        (p.mode == SRM_STORE_F16 || p.mode == SRM_ADD_STORE_F16 || p.mode == SRM_OUT_NCHW)) {
        const int st = (c.npad == 64) ? launch_conv_k64<64>(p, wk64, s) : (c.npad == 16) ? launch_conv_k64<16>(p, wk64, s) : K4_ERR_UNSUPPORTED;
        if (st != K4_ERR_UNSUPPORTED) return st;
    }
    if (c.npad == 64) return launch_conv_ws<64>(p, s);
    if (c.npad == 32) return launch_conv_ws<32>(p, s);
    return launch_conv_ws<16>(p, s);
}

// nearest-x2 + conv as four sub-pixel convolutions over the source grid (p.H x p.W = SOURCE size, dst is 2H x 2W)
int run_conv_up2x(const SrConv& c, ConvParams p, cudaStream_t s) {
    for (int ph = 0; ph < 4; ++ph) {
        const int py = ph >> 1, px = ph & 1;
        ConvParams q = p;
        q.upsample = 0; q.wpack = c.wphase[ph];
        q.ntaps = 4;
        for (int t = 0; t < 4; ++t) { q.tap_dy[t] = (unsigned char)(py + (t >> 1)); q.tap_dx[t] = (unsigned char)(px + (t & 1)); }
        q.out_s = 2; q.out_oy = py; q.out_ox = px;
        const int st = run_conv(c, q, s, c.wk64_phase[ph]);
        if (st != K4_OK) return st;
    }
    return K4_OK;
}

template <int COUT>
int launch_sft(const SftParams& p, cudaStream_t s) {
    constexpr int NW = 2 * (32 * 32 + 32) + 2 * (COUT * 32 + COUT);
    sft_kernel<COUT><<<(unsigned)((p.P + 127) / 128), 128, NW * 4, s>>>(p);
    K4_CUDA_TRY(cudaGetLastError());
    return K4_OK;
}

template <int COUT>
int launch_sft_tc(const SftTcParams& p, cudaStream_t s) {
    const SftBlob BL = sft_blob_layout(COUT);
    constexpr int per_sm_c = 4;                           // TMEM: 128 columns per CTA
    // ask for enough shared memory that no more than `per_sm` CTAs can share an SM: a CTA beyond the
    // TMEM budget would otherwise sit in tcgen05.alloc until a neighbour exits
    int smem = ((BL.total + 1023) & ~1023) + 8192 + 64;
    if (smem < (200 * 1024) / per_sm_c) smem = (200 * 1024) / per_sm_c;
    if (int st = sr_set_smem(sft_tc_kernel<COUT>, 9 + (COUT == 64 ? 0 : 1), smem)) return st;
    const int sms = sr_dev()->sms;
    const int per_sm = per_sm_c;
    if (p.n_tiles <= 0) return K4_OK;
    int grid = p.n_tiles < sms * per_sm ? p.n_tiles : sms * per_sm;
    K4_CUDA_TRY(sr_launch(sft_tc_kernel<COUT>, (unsigned)grid, 128, (size_t)smem, s, p));
    return K4_OK;
}

int run_sft(const SrSft& f, SftParams p, cudaStream_t s) {
    static const bool use_fp32 = getenv("K4_SFT_FP32") != nullptr;      // development switch: fp32 FFMA kernel
    if (use_fp32) {
        p.w = f.w;
        return f.cout == 64 ? launch_sft<64>(p, s) : launch_sft<32>(p, s);
    }
    SftTcParams q{};
    q.cond = p.cond; q.blob = f.blob; q.x_f = p.x_f; q.x_h = p.x_h; q.xh_cstride = p.xh_cstride; q.xh_c0 = p.xh_c0;
    q.dst_h = p.dst_h; q.dst_cstride = p.dst_cstride; q.dst_c0 = p.dst_c0; q.dst_f = p.dst_f; q.res_f = p.res_f;
    q.res_scale = p.res_scale; q.P = p.P; q.p0 = p.p0; q.n_tiles = (int)((p.P + 127) / 128);
    return f.cout == 64 ? launch_sft_tc<64>(q, s) : launch_sft_tc<32>(q, s);
}

}  // namespace

extern "C" int k4_srnet_destroy(k4_srnet* n) {
    if (!n) return K4_OK;
    for (int i = 0; i < n->n_allocs; ++i) cudaFree(n->allocs[i]);
    delete n;
    return K4_OK;
}

extern "C" int k4_srnet_create(const k4_srnet_desc* d, k4_stream_t stream, k4_srnet** out) {
    if (!d || !out || !d->h_params) return K4_ERR_INVALID_ARG;
    *out = nullptr;
    if (d->num_feat != 64 || d->num_grow_ch != 32 || d->num_cond != 1 || d->n_in_colors != 3 || d->scale != 4 ||
        d->num_block < 1 || d->num_block > 32)
        return K4_ERR_UNSUPPORTED;          // the shipped configuration (run_sr.py:1353); others are not built
    const int expect = 2 * (1 + 4 + d->num_block * (3 * 13 + 4) + 4 + 5);
    if (d->n_params != expect) return K4_ERR_INVALID_ARG;
    for (int i = 0; i < d->n_params; ++i) if (!d->h_params[i]) return K4_ERR_INVALID_ARG;
    int st = k4_device_check();
    if (st != K4_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    k4_srnet* n = new (std::nothrow) k4_srnet();
    if (!n) return K4_ERR_INVALID_ARG;
    memset(n, 0, sizeof(*n));
    n->num_feat = 64; n->num_block = d->num_block; n->num_grow = 32; n->num_cond = 1; n->n_in = 3; n->scale = 4;
    const float* const* P = d->h_params;
    int k = 0;
#define SR_TRY(x) do { st = (x); if (st != K4_OK) { k4_srnet_destroy(n); return st; } } while (0)
    SR_TRY(make_conv(n, n->conv_first, P[k], P[k + 1], 64, 3, s)); k += 2;
    {   // CondNet.0 (64x1x3x3), .2 (64x64), .4 (64x64), .6 (32x64)
        const size_t nw = 64 * 9 + 64 + 2 * (64 * 64 + 64) + 32 * 64 + 32;
        SR_TRY(sr_alloc(n, (void**)&n->condnet_w, nw * 4));
        float* dd = n->condnet_w;
        const size_t cnt[8] = {576, 64, 4096, 64, 4096, 64, 2048, 32};
        for (int q = 0; q < 8; ++q) { cudaMemcpyAsync(dd, P[k + q], cnt[q] * 4, cudaMemcpyDeviceToDevice, s); dd += cnt[q]; }
        k += 8;
        SR_TRY(sr_alloc(n, (void**)&n->condnet_mma, (size_t)cnm_layout().total));
        K4_CUDA_TRY(cudaMemsetAsync(n->condnet_mma, 0, cnm_layout().total, s));
        pack_condnet_mma_kernel<<<16, 256, 0, s>>>(n->condnet_w, n->condnet_mma);
        K4_CUDA_TRY(cudaGetLastError());
    }
    for (int i = 0; i < d->num_block; ++i) {
        for (int j = 0; j < 3; ++j) {
            for (int c = 0; c < 5; ++c) {
                const int cin = 64 + 32 * c, cout = (c == 4) ? 64 : 32;
                SR_TRY(make_conv(n, n->rdb_conv[i][j][c], P[k], P[k + 1], cout, cin, s)); k += 2;
            }
            SR_TRY(make_sft(n, n->rdb_sft[i][j][0], P + k, 64, s)); k += 8;
            SR_TRY(make_sft(n, n->rdb_sft[i][j][1], P + k, 32, s)); k += 8;
        }
        SR_TRY(make_sft(n, n->rrdb_sft[i], P + k, 64, s)); k += 8;
    }
    SR_TRY(make_sft(n, n->sftbody, P + k, 64, s)); k += 8;
    SR_TRY(make_conv(n, n->conv_body, P[k], P[k + 1], 64, 64, s)); SR_TRY(make_conv_k64(n, n->conv_body, P[k], 64, 64, false, s)); k += 2;
    SR_TRY(make_conv(n, n->conv_up1, P[k], P[k + 1], 64, 64, s)); SR_TRY(make_conv_phases(n, n->conv_up1, P[k], 64, 64, s)); SR_TRY(make_conv_k64(n, n->conv_up1, P[k], 64, 64, true, s)); k += 2;
    SR_TRY(make_conv(n, n->conv_up2, P[k], P[k + 1], 64, 64, s)); SR_TRY(make_conv_phases(n, n->conv_up2, P[k], 64, 64, s)); SR_TRY(make_conv_k64(n, n->conv_up2, P[k], 64, 64, true, s)); k += 2;
    SR_TRY(make_conv(n, n->conv_hr, P[k], P[k + 1], 64, 64, s)); SR_TRY(make_conv_k64(n, n->conv_hr, P[k], 64, 64, false, s)); k += 2;
    SR_TRY(make_conv(n, n->conv_last, P[k], P[k + 1], 3, 64, s)); SR_TRY(make_conv_k64(n, n->conv_last, P[k], 3, 64, false, s)); k += 2;
#undef SR_TRY
    *out = n;
    return K4_OK;
}

// workspace layout for an h x w tile (P = h*w LR pixels)
struct SrWs { size_t in16, cond32, cond16, feat, trunkA, trunkB, cat, cat2, sbody, bf, up1, up2, hr, total; };
static SrWs sr_ws(int h, int w) {
    const size_t P = (size_t)h * w;
    SrWs o; size_t off = 0;
    auto take = [&](size_t b) { size_t r = off; off += (b + 255) & ~(size_t)255; return r; };
    o.in16 = take(P * 32 * 2); o.cond32 = take(P * 32 * 4); o.feat = take(P * 64 * 4);
    o.trunkA = take(P * 64 * 4); o.trunkB = take(P * 64 * 4); o.cat = take(P * 192 * 2);
    o.cat2 = take(P * 192 * 2); o.cond16 = take(P * 32 * 2);         // fused SFT epilogues: dense blocks alternate between two concat buffers
    o.sbody = take(P * 64 * 2); o.bf = take(P * 64 * 2);
    o.up1 = take(P * 4 * 64 * 2); o.up2 = take(P * 16 * 64 * 2); o.hr = take(P * 16 * 64 * 2);
    o.total = off;
    return o;
}

extern "C" size_t k4_srnet_workspace_bytes(const k4_srnet*, int32_t h, int32_t w) {
    if (h <= 0 || w <= 0) return 0;
    return sr_ws(h, w).total;
}

// Rows [lo, hi) of an h-row tile that a layer must produce when `r` more rows of receptive field follow it
// (keep rows [ky0, ky1)); clipped to the tile -- beyond it the convolutions see their zero padding.
struct SrRows { int lo, hi; };
static inline SrRows sr_rows(int ky0, int ky1, int r, int h) {
    SrRows o; o.lo = ky0 - r < 0 ? 0 : ky0 - r; o.hi = ky1 + r > h ? h : ky1 + r; return o;
}

static int srnet_forward_roi_impl(const k4_srnet* n, const float* d_x, const float* d_cond, int32_t h, int32_t w,
                                  int32_t keep_y0, int32_t keep_y1, int32_t keep_x0, int32_t keep_x1,
                                  float* d_out, int64_t out_plane_stride, int64_t out_row_stride,
                                  int32_t n_extra, float* const* h_extra,
                                  void* d_ws, size_t ws_bytes, k4_stream_t stream) {
    if (!n || !d_x || !d_cond || !d_out || h <= 0 || w <= 0) return K4_ERR_INVALID_ARG;
    if (n_extra < 0 || n_extra > K4_MAX_PEERS - 1 || (n_extra > 0 && !h_extra)) return K4_ERR_INVALID_ARG;
    for (int e = 0; e < n_extra; ++e) if (!h_extra[e]) return K4_ERR_INVALID_ARG;
    if (keep_y0 < 0 || keep_y1 > h || keep_y0 >= keep_y1 || keep_x0 < 0 || keep_x1 > w || keep_x0 >= keep_x1) return K4_ERR_INVALID_ARG;
    const SrWs L = sr_ws(h, w);
    if (!d_ws || ws_bytes < L.total) return K4_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char* ws = (unsigned char*)d_ws;
    const long long P = (long long)h * w;
    __half* in16 = (__half*)(ws + L.in16); float* cond32 = (float*)(ws + L.cond32); float* feat = (float*)(ws + L.feat);
    float* tA = (float*)(ws + L.trunkA); float* tB = (float*)(ws + L.trunkB); __half* cat = (__half*)(ws + L.cat);
    __half* sbody = (__half*)(ws + L.sbody); __half* bf = (__half*)(ws + L.bf);
    __half* up1 = (__half*)(ws + L.up1); __half* up2 = (__half*)(ws + L.up2); __half* hr = (__half*)(ws + L.hr);
    __half* cat2 = (__half*)(ws + L.cat2); __half* cond16 = (__half*)(ws + L.cond16);
    // K4_SR_FUSE_SFT=0: every SFT layer as its own sft_tc_kernel pass (the round-1 structure, kept as the A/B reference)
    const char* fenv = getenv("K4_SR_FUSE_SFT");                   // read per call: tests compare both structures in one process
    const bool fuse = !(fenv && fenv[0] == '0') && !sr_use_v1() && !sr_no_tma();
    const int ky0 = keep_y0, ky1 = keep_y1;
    // Remaining receptive radius (LR rows) behind each layer, from the output backwards: conv_last / conv_hr / up2 / up1
    // need 1/4 + 1/4 + 1/4 + 1/2 LR rows (handled exactly in high-resolution rows below), conv_body 1, every 3x3 conv of a
    // residual dense block 1 (SFT layers 0), conv_first 1: the trunk entering conv_body's SFT is needed on rows(3), the
    // output of the k-th dense block from the end on rows(3 + 5k), conv_first's on rows(3 + 5 * 3 * num_block).
    const int nrdb = 3 * n->num_block;
    int st;
#define SR_DO(x) do { st = (x); if (st != K4_OK) return st; } while (0)
    K4_CUDA_TRY(sr_launch(nchw_to_nhwc32_kernel, (unsigned)((P * 32 + 255) / 256), 256, 0, s, d_x, in16, 3, P));
    {
        const SrRows rr = sr_rows(ky0, ky1, 3 + 5 * nrdb + 5, h);          // every SFT layer inside the window reads it
        const char* cenv = getenv("K4_CONDNET_FP32");
        if (cenv && cenv[0] == '1') {
            CondParams cp{d_cond, n->condnet_w, cond32, h, w, rr.lo, rr.hi, fuse ? cond16 : nullptr};
            constexpr int NW = 64 * 9 + 64 + 2 * (64 * 64 + 64) + 32 * 64 + 32;
            SR_DO(sr_set_smem(condnet_kernel, 11, NW * 4));
            const long long np = (long long)(rr.hi - rr.lo) * w;
            K4_CUDA_TRY(sr_launch(condnet_kernel, (unsigned)((np + 127) / 128), 128, (size_t)NW * 4, s, cp));
        } else {
            CondMmaParams cp{d_cond, n->condnet_mma, cond32, fuse ? cond16 : nullptr, h, w, rr.lo, rr.hi};
            const int smem = cnm_layout().total;
            SR_DO(sr_set_smem(condnet_mma_kernel, 12, smem));
            const long long nb = ((long long)(rr.hi - rr.lo) * w + 15) / 16;
            long long grid = (nb + 3) / 4;
            const long long cap = (long long)sr_dev()->sms * 4;
            if (grid > cap) grid = cap;
            K4_CUDA_TRY(sr_launch(condnet_mma_kernel, (unsigned)(grid < 1 ? 1 : grid), 128, (size_t)smem, s, cp));
        }
    }
    ConvParams c0{};
    c0.H = h; c0.W = w;
    auto rows = [&](ConvParams& p, int r) { const SrRows q = sr_rows(ky0, ky1, r, h); p.y_lo = q.lo; p.y_hi = q.hi; };
    auto win = [&](SftParams& sp, int r) { const SrRows q = sr_rows(ky0, ky1, r, h); sp.p0 = (long long)q.lo * w; sp.P = (long long)(q.hi - q.lo) * w; };
    if (fuse) {
        // Every SFT layer runs in the epilogue of the convolution that produces its input (epilogue_sft): conv_first -> sft0 of
        // the first dense block, conv4 -> sft1, conv5 -> sft0 of the next block (for the third block of an RRDB: the RRDB
        // tail first, then the next sft0 or sftbody).  A block's conv5 writes the NEXT block's first 64 concat channels while
        // other CTAs still read this block's, so the dense blocks alternate between two concat buffers.
        auto sft_of = [&](int r) -> const SrSft& {                 // sft0 of dense block r (global index), or sftbody after the last
            return r < nrdb ? n->rdb_sft[r / 3][r % 3][0] : n->sftbody;
        };
        auto next_dst = [&](ConvParams& p, int r) {                // where the fused sft0 / sftbody result goes
            if (r < nrdb) { p.dst_h2 = (r & 1) ? cat2 : cat; p.dst2_cstride = 192; p.dst2_c0 = 0; }
            else { p.dst_h2 = sbody; p.dst2_cstride = 64; p.dst2_c0 = 0; }
        };
        {   // feat = conv_first(x) -> feat, trunk A; xc0 = sft0(feat) -> cat[0:64]
            ConvParams p = c0; p.src = in16; p.src_cstride = 32; p.src_c0 = 0; p.mode = SRM_TRUNK_SFT; p.scale = 1.f;
            p.dst_f = feat; p.dst_f2 = tA; p.cond16 = cond16; p.sftw = sft_of(0).frag; p.sft_n = 64; next_dst(p, 0);
            rows(p, 3 + 5 * nrdb);
            SR_DO(run_conv(n->conv_first, p, s));
        }
        for (int i = 0; i < n->num_block; ++i) {
            const float* cur = tA;
            for (int j = 0; j < 3; ++j) {
                const int r = 3 * i + j, k = nrdb - 1 - r, base = 3 + 5 * k;
                __half* cc = (r & 1) ? cat2 : cat;
                for (int c = 0; c < 4; ++c) {   // x{c+1} = lrelu(conv(cat[0:64+32c])) -> cat[64+32c : 96+32c]; conv4 also applies sft1
                    ConvParams p = c0; p.src = cc; p.src_cstride = 192; p.src_c0 = 0; p.mode = SRM_STORE_F16; p.lrelu = 0.2f;
                    p.dst_h = cc; p.dst_cstride = 192; p.dst_c0 = 64 + 32 * c;
                    if (c == 3) { p.mode = SRM_STORE_F16_SFT; p.cond16 = cond16; p.sftw = n->rdb_sft[i][j][1].frag; p.sft_n = 32; }
                    rows(p, base + 4 - c);
                    SR_DO(run_conv(n->rdb_conv[i][j][c], p, s));
                }
                {   // x = conv5(cat) * 0.2 + x, then the SFT layer(s) that read it
                    ConvParams p = c0; p.src = cc; p.src_cstride = 192; p.src_c0 = 0; p.scale = 0.2f; p.add_f = cur;
                    p.cond16 = cond16; p.sft_n = 64;
                    if (j < 2) {
                        p.mode = SRM_TRUNK_SFT; p.dst_f = tB; p.sftw = sft_of(r + 1).frag;
                    } else {            // block tail: out = sft0(x) * 0.2 + x_in -> trunk A, then the next sft0 / sftbody of that
                        p.mode = SRM_TRUNK_SFT2; p.dst_f = tA; p.add_f2 = tA; p.scale2 = 0.2f;
                        p.sftw = n->rrdb_sft[i].frag; p.sftw2 = sft_of(r + 1).frag;
                    }
                    next_dst(p, r + 1);
                    rows(p, base);
                    SR_DO(run_conv(n->rdb_conv[i][j][4], p, s));
                    cur = tB;
                }
            }
        }
        {   // body_feat = conv_body(sftbody(trunk)) + feat   (sbody was written by the last conv5)
            ConvParams p = c0; p.src = sbody; p.src_cstride = 64; p.mode = SRM_ADD_STORE_F16; p.add_f = feat; p.dst_h = bf; p.dst_cstride = 64;
            rows(p, 2);
            SR_DO(run_conv(n->conv_body, p, s));
        }
    } else {
        {   // feat = conv_first(x): fp32 into `feat` and into trunk A (the initial trunk)
            ConvParams p = c0; p.src = in16; p.src_cstride = 32; p.src_c0 = 0; p.mode = SRM_STORE_F32F16; p.dst_f = feat; p.dst_f2 = tA;
            rows(p, 3 + 5 * nrdb);
            SR_DO(run_conv(n->conv_first, p, s));
        }
        for (int i = 0; i < n->num_block; ++i) {
            // RRDB_SFT.forward: trunk A holds x (kept for the block's tail), RDBs update A -> B -> B -> B
            const float* cur = tA;
            for (int j = 0; j < 3; ++j) {
                const int k = nrdb - 1 - (3 * i + j);                              // dense blocks still to come
                const int base = 3 + 5 * k;                                        // rows(base): this block's output
                {   // xc0 = sft0(x) -> cat[0:64]
                    SftParams sp{}; sp.cond = cond32; sp.x_f = cur; sp.dst_h = cat; sp.dst_cstride = 192; sp.dst_c0 = 0;
                    win(sp, base + 5);
                    SR_DO(run_sft(n->rdb_sft[i][j][0], sp, s));
                }
                for (int c = 0; c < 4; ++c) {   // x{c+1} = lrelu(conv(cat[0:64+32c])) -> cat[64+32c : 96+32c]
                    ConvParams p = c0; p.src = cat; p.src_cstride = 192; p.src_c0 = 0; p.mode = SRM_STORE_F16; p.lrelu = 0.2f;
                    p.dst_h = cat; p.dst_cstride = 192; p.dst_c0 = 64 + 32 * c;
                    rows(p, base + 4 - c);
                    SR_DO(run_conv(n->rdb_conv[i][j][c], p, s));
                }
                {   // xc1 = sft1(x4) in place: cat[160:192]
                    SftParams sp{}; sp.cond = cond32; sp.x_h = cat; sp.xh_cstride = 192; sp.xh_c0 = 160;
                    sp.dst_h = cat; sp.dst_cstride = 192; sp.dst_c0 = 160;
                    win(sp, base + 1);
                    SR_DO(run_sft(n->rdb_sft[i][j][1], sp, s));
                }
                {   // x = conv5(cat) * 0.2 + x
                    ConvParams p = c0; p.src = cat; p.src_cstride = 192; p.src_c0 = 0; p.mode = SRM_TRUNK; p.scale = 0.2f;
                    p.add_f = cur; p.dst_f = tB;
                    rows(p, base);
                    SR_DO(run_conv(n->rdb_conv[i][j][4], p, s));
                    cur = tB;
                }
            }
            {   // out = sft0(rdb3 out) * 0.2 + x_in  -> A
                SftParams sp{}; sp.cond = cond32; sp.x_f = tB; sp.dst_f = tA; sp.res_f = tA; sp.res_scale = 0.2f;
                win(sp, 3 + 5 * (nrdb - 3 * (i + 1)));
                SR_DO(run_sft(n->rrdb_sft[i], sp, s));
            }
        }
        {   // body_feat = conv_body(sftbody(trunk)) + feat
            SftParams sp{}; sp.cond = cond32; sp.x_f = tA; sp.dst_h = sbody; sp.dst_cstride = 64; sp.dst_c0 = 0;
            win(sp, 3);
            SR_DO(run_sft(n->sftbody, sp, s));
            ConvParams p = c0; p.src = sbody; p.src_cstride = 64; p.mode = SRM_ADD_STORE_F16; p.add_f = feat; p.dst_h = bf; p.dst_cstride = 64;
            rows(p, 2);
            SR_DO(run_conv(n->conv_body, p, s));
        }
    }
    {   // upsample x2 (nearest) + conv + lrelu, twice; conv_hr + lrelu; conv_last (rows in the layer's own resolution)
        ConvParams p = c0; p.H = 2 * h; p.W = 2 * w; p.upsample = 1; p.src = bf; p.src_cstride = 64; p.mode = SRM_STORE_F16; p.lrelu = 0.2f;
        p.dst_h = up1; p.dst_cstride = 64;
        auto clip = [](int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); };
        if (sr_use_v1()) {
            p.y_lo = clip(2 * ky0 - 2, 2 * h); p.y_hi = clip(2 * ky1 + 2, 2 * h);
            SR_DO(run_conv(n->conv_up1, p, s));
            p.H = 4 * h; p.W = 4 * w; p.src = up1; p.dst_h = up2;
            p.y_lo = clip(4 * ky0 - 2, 4 * h); p.y_hi = clip(4 * ky1 + 2, 4 * h);
            SR_DO(run_conv(n->conv_up2, p, s));
        } else {
            p.H = h; p.W = w;                                   // iterate over the SOURCE grid, write the 2x grid
            p.y_lo = clip(ky0 - 1, h); p.y_hi = clip(ky1 + 1, h);                    // -> up1 rows [2ky0-2, 2ky1+2)
            SR_DO(run_conv_up2x(n->conv_up1, p, s));
            p.H = 2 * h; p.W = 2 * w; p.src = up1; p.dst_h = up2;
            p.y_lo = clip(2 * ky0 - 1, 2 * h); p.y_hi = clip(2 * ky1 + 1, 2 * h);    // -> up2 rows [4ky0-2, 4ky1+2)
            SR_DO(run_conv_up2x(n->conv_up2, p, s));
            p.H = 4 * h; p.W = 4 * w;
        }
        p.upsample = 0; p.src = up2; p.dst_h = hr;
        p.y_lo = clip(4 * ky0 - 1, 4 * h); p.y_hi = clip(4 * ky1 + 1, 4 * h);
        SR_DO(run_conv(n->conv_hr, p, s));
        ConvParams q = c0; q.H = 4 * h; q.W = 4 * w; q.src = hr; q.src_cstride = 64; q.mode = SRM_OUT_NCHW; q.out_nchw = d_out; q.n_valid = 3;
        q.y_lo = 4 * ky0; q.y_hi = 4 * ky1;
        q.out_ps = out_plane_stride; q.out_rs = out_row_stride;
        q.crop_y0 = 4 * ky0; q.crop_y1 = 4 * ky1; q.crop_x0 = 4 * keep_x0; q.crop_x1 = 4 * keep_x1;
        q.n_out_extra = n_extra;
        for (int e = 0; e < n_extra; ++e) q.out_extra[e] = h_extra[e];
        SR_DO(run_conv(n->conv_last, q, s));
    }
#undef SR_DO
    return K4_OK;
}

extern "C" int k4_srnet_forward_roi(const k4_srnet* n, const float* d_x, const float* d_cond, int32_t h, int32_t w,
                                    int32_t keep_y0, int32_t keep_y1, int32_t keep_x0, int32_t keep_x1,
                                    float* d_out, int64_t out_plane_stride, int64_t out_row_stride,
                                    void* d_ws, size_t ws_bytes, k4_stream_t stream) {
    return srnet_forward_roi_impl(n, d_x, d_cond, h, w, keep_y0, keep_y1, keep_x0, keep_x1, d_out, out_plane_stride,
                                  out_row_stride, 0, nullptr, d_ws, ws_bytes, stream);
}

extern "C" int k4_srnet_forward_roi_peers(const k4_srnet* n, const float* d_x, const float* d_cond, int32_t h, int32_t w,
                                          int32_t keep_y0, int32_t keep_y1, int32_t keep_x0, int32_t keep_x1,
                                          float* d_out, int64_t out_plane_stride, int64_t out_row_stride,
                                          int32_t n_extra, float* const* h_extra,
                                          void* d_ws, size_t ws_bytes, k4_stream_t stream) {
    return srnet_forward_roi_impl(n, d_x, d_cond, h, w, keep_y0, keep_y1, keep_x0, keep_x1, d_out, out_plane_stride,
                                  out_row_stride, n_extra, h_extra, d_ws, ws_bytes, stream);
}

extern "C" int k4_srnet_forward(const k4_srnet* n, const float* d_x, const float* d_cond, int32_t h, int32_t w,
                                float* d_out, void* d_ws, size_t ws_bytes, k4_stream_t stream) {
    return k4_srnet_forward_roi(n, d_x, d_cond, h, w, 0, h, 0, w, d_out, (int64_t)16 * h * w, (int64_t)4 * w, d_ws, ws_bytes, stream);
}

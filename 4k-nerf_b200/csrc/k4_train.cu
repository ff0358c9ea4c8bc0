// k4_train.cu -- the grid-maintenance and optimiser steps that sit either side of the render path
// during training.  All are HBM-bound element-wise / stencil passes over
// the voxel grids; one launch each, on the caller's stream, no host synchronisation.
//
//   reference                                                   entry point
//   total_variation_add_grad   lib/cuda/total_variation_kernel.cu:13-66      k4_op_total_variation_add_grad
//   adam_upd / masked_adam_upd / adam_upd_with_perlr
//                              lib/cuda/adam_upd_kernel.cu:8-136             k4_op_adam_upd
//   update_occupancy_cache     lib/dvgo.py:224-233, lib/dmpigo.py:212-224    k4_op_grid_alpha + k4_op_maxpool3_thres_and
//   DenseGrid.scale_volume_grid lib/grid.py:130-135 (F.interpolate trilinear, align_corners=True)
//                                                                            k4_op_resample_trilinear
//
// The first two are bit-identical to the reference extension (floating-point shape pinned to the
// reference build's SASS: see the comments at each expression; tests/test_gpu_train_ops.py compares
// against oracle/_ref).  The last two restate ATen kernels and are held to fp32 rounding.
#include <cstring>
#include "k4_march_common.cuh"

namespace {

constexpr int TR_T = 256;
inline unsigned tr_blocks(long long n) { return (unsigned)((n + TR_T - 1) / TR_T); }

__device__ __forceinline__ float clamp1(float v) { return fminf(fmaxf(v, -1.f), 1.f); }

// One thread per grid element.  `param` is [1,C,I,J,K] contiguous; the channel is folded into the
// leading index exactly as the reference does (i = idx / K / J % I), so channels never mix.
template <bool DENSE>
__global__ void tv_add_grad_kernel(const float* __restrict__ param, float* __restrict__ grad, float wx, float wy, float wz,
                                   long long szi, long long szj, long long szk, long long n) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const float g0 = grad[idx];
    if (!DENSE && g0 == 0.f) return;
    const long long k = idx % szk;
    const long long j = idx / szk % szj;
    const long long i = idx / szk / szj % szi;
    const float p = __ldg(param + idx);
    float acc = 0.f;
    // reference build: every term is FMUL w, clamp then FADD into the running sum (the select on the
    // boundary keeps ptxas from contracting), first term added to +0
    acc = __fadd_rn(acc, (k == 0) ? 0.f : __fmul_rn(wx, clamp1(__fsub_rn(p, __ldg(param + idx - 1)))));
    acc = __fadd_rn(acc, (k == szk - 1) ? 0.f : __fmul_rn(wx, clamp1(__fsub_rn(p, __ldg(param + idx + 1)))));
    acc = __fadd_rn(acc, (j == 0) ? 0.f : __fmul_rn(wy, clamp1(__fsub_rn(p, __ldg(param + idx - szk)))));
    acc = __fadd_rn(acc, (j == szj - 1) ? 0.f : __fmul_rn(wy, clamp1(__fsub_rn(p, __ldg(param + idx + szk)))));
    acc = __fadd_rn(acc, (i == 0) ? 0.f : __fmul_rn(wz, clamp1(__fsub_rn(p, __ldg(param + idx - szk * szj)))));
    acc = __fadd_rn(acc, (i == szi - 1) ? 0.f : __fmul_rn(wz, clamp1(__fsub_rn(p, __ldg(param + idx + szk * szj)))));
    grad[idx] = __fadd_rn(g0, acc);
}

// Adam moment + parameter update of one element (adam_upd_kernel.cu:19-23 and twins).
__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float lr_scale, float step_size,
                                          float beta1, float beta2, float omb1, float omb2, float eps, bool perlr) {
    m = __fmaf_rn(beta1, m, __fmul_rn(omb1, g));
    v = __fmaf_rn(beta2, v, __fmul_rn(__fmul_rn(omb2, g), g));
    float num = perlr ? __fmul_rn(__fmul_rn(step_size, lr_scale), m) : __fmul_rn(step_size, m);
    p = __fsub_rn(p, __fdiv_rn(num, __fadd_rn(__fsqrt_rn(v), eps)));
}

// Four elements per thread (float4) when the arrays allow it: with skip_zero_grad a thread whose four
// gradients are all zero stops after the 16-byte gradient load, so sparse-gradient steps (the usual
// case: only voxels touched by the batch's rays have gradients) move 4 B/element instead of 28.
template <bool MASKED, bool PERLR, int VEC>
__global__ void adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                            const float* __restrict__ perlr, long long n, float step_size, float beta1, float beta2, float eps) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const float omb1 = __fsub_rn(1.f, beta1), omb2 = __fsub_rn(1.f, beta2);
    if (VEC == 4) {
        if (t * 4 >= n) return;
        const float4 g = __ldg(reinterpret_cast<const float4*>(grad) + t);
        if (MASKED && g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) return;
        float4 p = reinterpret_cast<float4*>(param)[t], mm = reinterpret_cast<float4*>(m)[t], vv = reinterpret_cast<float4*>(v)[t];
        float4 l = make_float4(1.f, 1.f, 1.f, 1.f);
        if (PERLR) l = __ldg(reinterpret_cast<const float4*>(perlr) + t);
        if (!MASKED || g.x != 0.f) adam_elem(p.x, g.x, mm.x, vv.x, l.x, step_size, beta1, beta2, omb1, omb2, eps, PERLR);
        if (!MASKED || g.y != 0.f) adam_elem(p.y, g.y, mm.y, vv.y, l.y, step_size, beta1, beta2, omb1, omb2, eps, PERLR);
        if (!MASKED || g.z != 0.f) adam_elem(p.z, g.z, mm.z, vv.z, l.z, step_size, beta1, beta2, omb1, omb2, eps, PERLR);
        if (!MASKED || g.w != 0.f) adam_elem(p.w, g.w, mm.w, vv.w, l.w, step_size, beta1, beta2, omb1, omb2, eps, PERLR);
        reinterpret_cast<float4*>(param)[t] = p;
        reinterpret_cast<float4*>(m)[t] = mm;
        reinterpret_cast<float4*>(v)[t] = vv;
    } else {
        if (t >= n) return;
        const float g = grad[t];
        if (MASKED && g == 0.f) return;
        float p = param[t], mm = m[t], vv = v[t];
        adam_elem(p, g, mm, vv, PERLR ? perlr[t] : 1.f, step_size, beta1, beta2, omb1, omb2, eps, PERLR);
        param[t] = p; m[t] = mm; v[t] = vv;
    }
}

// alpha at every point of the occupancy grid: trilinear density (DenseGrid.forward, the marcher's own
// gather) at (lx[i], ly[j], lz[k]) -> 1 - (1 + exp(d + shift))^(-interval)  (Raw2Alpha, lib/dvgo.py:453-466)
__global__ void grid_alpha_kernel(K4Dev s, const float* __restrict__ lx, const float* __restrict__ ly, const float* __restrict__ lz,
                                  int mX, int mY, int mZ, float shift, float interval, float* __restrict__ alpha) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)mX * mY * mZ) return;
    const int k = (int)(idx % mZ), j = (int)(idx / mZ % mY), i = (int)(idx / mZ / mY);
    Cell cell = make_cell(s, __ldg(lx + i), __ldg(ly + j), __ldg(lz + k));
    float cw[8];
    int cidx[8];
    corner_setup(s, cell, cw, cidx);
    const float den = interp_density(s, cw, cidx);
    const float e = expf(__fadd_rn(den, shift));
    alpha[idx] = __fsub_rn(1.f, powf(__fadd_rn(1.f, e), -interval));
}

// mask &= maxpool3x3x3(alpha, stride 1, pad 1) > thres   (F.max_pool3d pads with -inf)
__global__ void maxpool3_thres_and_kernel(const float* __restrict__ alpha, int mX, int mY, int mZ, float thres, uint8_t* __restrict__ mask) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)mX * mY * mZ) return;
    if (!mask[idx]) return;                                   // already empty: nothing to decide, nothing to read
    const int k = (int)(idx % mZ), j = (int)(idx / mZ % mY), i = (int)(idx / mZ / mY);
    float mx = -INFINITY;
    for (int di = -1; di <= 1; ++di) {
        const int ii = i + di;
        if (ii < 0 || ii >= mX) continue;
        for (int dj = -1; dj <= 1; ++dj) {
            const int jj = j + dj;
            if (jj < 0 || jj >= mY) continue;
            const float* row = alpha + ((long long)ii * mY + jj) * mZ;
#pragma unroll
            for (int dk = -1; dk <= 1; ++dk) {
                const int kk = k + dk;
                if (kk >= 0 && kk < mZ) mx = fmaxf(mx, __ldg(row + kk));
            }
        }
    }
    mask[idx] = (uint8_t)(mx > thres);
}

// ATen upsample_trilinear3d, align_corners=True: src = dst * (in-1)/(out-1) (fp32), lambda = frac.
__device__ __forceinline__ void up_axis(int o, int in, int out, int& i0, int& ip, float& l0, float& l1) {
    const float r = (out > 1) ? __fdiv_rn((float)(in - 1), (float)(out - 1)) : 0.f;
    const float sr = __fmul_rn(r, (float)o);
    i0 = (int)sr;
    ip = (i0 < in - 1) ? 1 : 0;
    l1 = __fsub_rn(sr, (float)i0);
    l0 = __fsub_rn(1.f, l1);
}

__global__ void resample_trilinear_kernel(const float* __restrict__ src, int C, int X, int Y, int Z, float* __restrict__ dst,
                                          int X2, int Y2, int Z2) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n2 = (long long)X2 * Y2 * Z2;
    if (idx >= n2) return;
    const int z = (int)(idx % Z2), y = (int)(idx / Z2 % Y2), x = (int)(idx / Z2 / Y2);
    int x0, xp, y0, yp, z0, zp;
    float a0, a1, b0, b1, c0, c1;
    up_axis(x, X, X2, x0, xp, a0, a1);
    up_axis(y, Y, Y2, y0, yp, b0, b1);
    up_axis(z, Z, Z2, z0, zp, c0, c1);
    const long long n1 = (long long)X * Y * Z;
    const long long o000 = ((long long)x0 * Y + y0) * Z + z0;
    const long long dxs = (long long)xp * Y * Z, dys = (long long)yp * Z, dzs = zp;
    for (int c = 0; c < C; ++c) {
        const float* p = src + (long long)c * n1 + o000;
        const float v00 = __fmaf_rn(c1, __ldg(p + dzs), __fmul_rn(c0, __ldg(p)));
        const float v01 = __fmaf_rn(c1, __ldg(p + dys + dzs), __fmul_rn(c0, __ldg(p + dys)));
        const float v10 = __fmaf_rn(c1, __ldg(p + dxs + dzs), __fmul_rn(c0, __ldg(p + dxs)));
        const float v11 = __fmaf_rn(c1, __ldg(p + dxs + dys + dzs), __fmul_rn(c0, __ldg(p + dxs + dys)));
        const float u0 = __fmaf_rn(b1, v01, __fmul_rn(b0, v00));
        const float u1 = __fmaf_rn(b1, v11, __fmul_rn(b0, v10));
        dst[(long long)c * n2 + idx] = __fmaf_rn(a1, u1, __fmul_rn(a0, u0));
    }
}

// ub360_utils_kernel.cu:12-32: per ray, running sum of consecutive-sample distances; emit and reset when it
// exceeds thres (the reset is a multiply by 0/1 in the reference: same values for finite sums).
__global__ void cumdist_thres_kernel(const float* __restrict__ dist, float thres, long long n_rays, long long n_pts, uint8_t* __restrict__ mask) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    float cum = 0.f;
    for (long long i = r * n_pts; i < (r + 1) * n_pts; ++i) {
        cum = __fadd_rn(cum, dist[i]);
        const bool over = cum > thres;
        cum = __fmul_rn(cum, over ? 0.f : 1.f);
        mask[i] = (uint8_t)over;
    }
}

#define TR_CHECK_LAUNCH()                                  \
    do {                                                   \
        cudaError_t e_ = cudaGetLastError();               \
        if (e_ != cudaSuccess) { k4_set_cuda_error(e_, "k4_train launch"); return K4_ERR_CUDA; } \
        return K4_OK;                                      \
    } while (0)


// ---------------------------------------------------------------------------------------------
// DenseGrid.forward with autograd (lib/grid.py:117-128), the training-side twin of the marcher's
// in-kernel interpolation: ((xyz - min) / (max - min)).flip(-1) * 2 - 1, ATen grid_sampler_3d
// (bilinear, zeros padding, align_corners=True) and the [C,M] -> [M,C] transpose in ONE launch
// instead of the reference's ~6 element-wise kernels + grid_sampler + 2 layout copies; its backward
// scatters the output gradient into the grid (the reference gets it from ATen's
// grid_sampler_3d_backward, which also computes the unused gradient w.r.t. the sample positions).
// One thread per (point, channel); the grid keeps the reference's planar [1,C,X,Y,Z] layout, it is the
// nn.Parameter the optimiser steps.  Corner order and FMA chain as interp_density (k4_march_common.cuh).
// ---------------------------------------------------------------------------------------------
struct GsParams { K4Dev s; const float* grid; const float* xyz; long long M; int C; long long nvox; };

__global__ void grid_sample_fwd_kernel(const __grid_constant__ GsParams p, float* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.M * p.C) return;
    const long long m = t / p.C;
    const int c = (int)(t - m * p.C);
    const Cell cell = make_cell(p.s, __ldg(p.xyz + 3 * m), __ldg(p.xyz + 3 * m + 1), __ldg(p.xyz + 3 * m + 2));
    float w[8]; int idx[8];
    corner_setup(p.s, cell, w, idx);
    const float* g = p.grid + (long long)c * p.nvox;
    float acc = __fmul_rn(__ldg(g + idx[0]), w[0]);
#pragma unroll
    for (int k = 1; k < 8; ++k) acc = __fmaf_rn(__ldg(g + idx[k]), w[k], acc);
    out[t] = acc;
}

__global__ void grid_sample_bwd_kernel(const __grid_constant__ GsParams p, const float* __restrict__ grad_out, float* __restrict__ grad_grid) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.M * p.C) return;
    const float go = __ldg(grad_out + t);
    if (go == 0.f) return;
    const long long m = t / p.C;
    const int c = (int)(t - m * p.C);
    const Cell cell = make_cell(p.s, __ldg(p.xyz + 3 * m), __ldg(p.xyz + 3 * m + 1), __ldg(p.xyz + 3 * m + 2));
    float w[8]; int idx[8];
    corner_setup(p.s, cell, w, idx);
    float* g = grad_grid + (long long)c * p.nvox;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (w[k] != 0.f) atomicAdd(g + idx[k], __fmul_rn(w[k], go));       // RED.ADD.F32: no return value, resolved in L2
}

}  // namespace

extern "C" {

int k4_op_total_variation_add_grad(const float* d_param, float* d_grad, float wx, float wy, float wz, int32_t dense_mode,
                                   int64_t n, int32_t sz_i, int32_t sz_j, int32_t sz_k, k4_stream_t stream) {
    if (n < 0 || sz_i <= 0 || sz_j <= 0 || sz_k <= 0) return K4_ERR_INVALID_ARG;
    if (n == 0) return K4_OK;
    if (!d_param || !d_grad) return K4_ERR_INVALID_ARG;
    wx /= 6; wy /= 6; wz /= 6;                                  // total_variation_kernel.cu:45-47 (fp32)
    cudaStream_t s = (cudaStream_t)stream;
    if (dense_mode) tv_add_grad_kernel<true><<<tr_blocks(n), TR_T, 0, s>>>(d_param, d_grad, wx, wy, wz, sz_i, sz_j, sz_k, n);
    else tv_add_grad_kernel<false><<<tr_blocks(n), TR_T, 0, s>>>(d_param, d_grad, wx, wy, wz, sz_i, sz_j, sz_k, n);
    TR_CHECK_LAUNCH();
}

int k4_op_adam_upd(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, const float* d_perlr, int64_t n,
                   int32_t step, float beta1, float beta2, float lr, float eps, int32_t skip_zero_grad, k4_stream_t stream) {
    if (n < 0 || step < 1) return K4_ERR_INVALID_ARG;
    if (n == 0) return K4_OK;
    if (!d_param || !d_grad || !d_exp_avg || !d_exp_avg_sq) return K4_ERR_INVALID_ARG;
    // adam_upd_kernel.cu:76 -- fp32 throughout (float overloads of pow / sqrt on the host)
    const float step_size = lr * sqrtf(1 - powf(beta2, (float)step)) / (1 - powf(beta1, (float)step));
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = (n % 4 == 0) && (((uintptr_t)d_param | (uintptr_t)d_grad | (uintptr_t)d_exp_avg | (uintptr_t)d_exp_avg_sq |
                                       (uintptr_t)d_perlr) % 16 == 0);
    const long long nt = vec ? n / 4 : n;
#define K4_ADAM(M, P)                                                                                              \
    do {                                                                                                           \
        if (vec) adam_kernel<M, P, 4><<<tr_blocks(nt), TR_T, 0, s>>>(d_param, d_grad, d_exp_avg, d_exp_avg_sq, d_perlr, n, step_size, beta1, beta2, eps); \
        else adam_kernel<M, P, 1><<<tr_blocks(nt), TR_T, 0, s>>>(d_param, d_grad, d_exp_avg, d_exp_avg_sq, d_perlr, n, step_size, beta1, beta2, eps);     \
    } while (0)
    if (d_perlr) K4_ADAM(false, true);                           // lib/masked_adam.py:60-63: per-voxel lr wins
    else if (skip_zero_grad) K4_ADAM(true, false);
    else K4_ADAM(false, false);
#undef K4_ADAM
    TR_CHECK_LAUNCH();
}

int k4_op_grid_alpha(const float* d_density, int32_t X, int32_t Y, int32_t Z, const float* h_xyz_min, const float* h_xyz_max,
                     const float* d_lx, const float* d_ly, const float* d_lz, int32_t mX, int32_t mY, int32_t mZ,
                     float shift, float interval, float* d_alpha, k4_stream_t stream) {
    if (!d_density || !h_xyz_min || !h_xyz_max || !d_lx || !d_ly || !d_lz || !d_alpha) return K4_ERR_INVALID_ARG;
    if (X <= 0 || Y <= 0 || Z <= 0 || mX <= 0 || mY <= 0 || mZ <= 0) return K4_ERR_INVALID_ARG;
    K4Dev v;
    memset(&v, 0, sizeof(v));
    v.X = X; v.Y = Y; v.Z = Z;
    for (int a = 0; a < 3; ++a) {
        v.xyz_min[a] = h_xyz_min[a]; v.xyz_max[a] = h_xyz_max[a];
        v.xyz_len[a] = h_xyz_max[a] - h_xyz_min[a];
    }
    v.density = d_density;
    const long long n = (long long)mX * mY * mZ;
    grid_alpha_kernel<<<tr_blocks(n), TR_T, 0, (cudaStream_t)stream>>>(v, d_lx, d_ly, d_lz, mX, mY, mZ, shift, interval, d_alpha);
    TR_CHECK_LAUNCH();
}

int k4_op_maxpool3_thres_and(const float* d_alpha, int32_t mX, int32_t mY, int32_t mZ, float thres, uint8_t* d_mask, k4_stream_t stream) {
    if (!d_alpha || !d_mask || mX <= 0 || mY <= 0 || mZ <= 0) return K4_ERR_INVALID_ARG;
    const long long n = (long long)mX * mY * mZ;
    maxpool3_thres_and_kernel<<<tr_blocks(n), TR_T, 0, (cudaStream_t)stream>>>(d_alpha, mX, mY, mZ, thres, d_mask);
    TR_CHECK_LAUNCH();
}

int k4_op_cumdist_thres(const float* d_dist, float thres, int64_t n_rays, int64_t n_pts, uint8_t* d_mask, k4_stream_t stream) {
    if (n_rays < 0 || n_pts < 0) return K4_ERR_INVALID_ARG;
    if (n_rays == 0 || n_pts == 0) return K4_OK;
    if (!d_dist || !d_mask) return K4_ERR_INVALID_ARG;
    cumdist_thres_kernel<<<tr_blocks(n_rays), TR_T, 0, (cudaStream_t)stream>>>(d_dist, thres, n_rays, n_pts, d_mask);
    TR_CHECK_LAUNCH();
}

int k4_op_resample_trilinear(const float* d_src, int32_t C, int32_t X, int32_t Y, int32_t Z, float* d_dst, int32_t X2, int32_t Y2,
                             int32_t Z2, k4_stream_t stream) {
    if (C < 0 || X <= 0 || Y <= 0 || Z <= 0 || X2 <= 0 || Y2 <= 0 || Z2 <= 0) return K4_ERR_INVALID_ARG;
    if (C == 0) return K4_OK;
    if (!d_src || !d_dst) return K4_ERR_INVALID_ARG;
    const long long n = (long long)X2 * Y2 * Z2;
    resample_trilinear_kernel<<<tr_blocks(n), TR_T, 0, (cudaStream_t)stream>>>(d_src, C, X, Y, Z, d_dst, X2, Y2, Z2);
    TR_CHECK_LAUNCH();
}


static int gs_params(GsParams& p, const float* d_grid, int32_t C, int32_t X, int32_t Y, int32_t Z, const float* h_xyz_min,
                     const float* h_xyz_max, const float* d_xyz, int64_t M) {
    if (!d_grid || !h_xyz_min || !h_xyz_max || C <= 0 || X < 1 || Y < 1 || Z < 1 || M < 0) return K4_ERR_INVALID_ARG;
    if (M > 0 && !d_xyz) return K4_ERR_INVALID_ARG;
    if ((long long)X * Y * Z >= (1ll << 31)) return K4_ERR_UNSUPPORTED;
    memset(&p, 0, sizeof(p));
    p.s.X = X; p.s.Y = Y; p.s.Z = Z;
    for (int a = 0; a < 3; ++a) {
        p.s.xyz_min[a] = h_xyz_min[a]; p.s.xyz_max[a] = h_xyz_max[a];
        p.s.xyz_len[a] = h_xyz_max[a] - h_xyz_min[a];                        // fp32, as torch: (xyz_max - xyz_min)
    }
    p.grid = d_grid; p.xyz = d_xyz; p.M = M; p.C = C; p.nvox = (long long)X * Y * Z;
    return K4_OK;
}

int k4_op_grid_sample(const float* d_grid, int32_t C, int32_t X, int32_t Y, int32_t Z, const float* h_xyz_min, const float* h_xyz_max,
                      const float* d_xyz, int64_t M, float* d_out, k4_stream_t stream) {
    GsParams p;
    int st = gs_params(p, d_grid, C, X, Y, Z, h_xyz_min, h_xyz_max, d_xyz, M);
    if (st != K4_OK) return st;
    if (M == 0) return K4_OK;
    if (!d_out) return K4_ERR_INVALID_ARG;
    grid_sample_fwd_kernel<<<tr_blocks(M * C), TR_T, 0, (cudaStream_t)stream>>>(p, d_out);
    TR_CHECK_LAUNCH();
}

int k4_op_grid_sample_backward(const float* d_grad_out, int32_t C, int32_t X, int32_t Y, int32_t Z, const float* h_xyz_min,
                               const float* h_xyz_max, const float* d_xyz, int64_t M, float* d_grad_grid, k4_stream_t stream) {
    GsParams p;
    int st = gs_params(p, d_grad_grid, C, X, Y, Z, h_xyz_min, h_xyz_max, d_xyz, M);
    if (st != K4_OK) return st;
    if (M == 0) return K4_OK;
    if (!d_grad_out) return K4_ERR_INVALID_ARG;
    grid_sample_bwd_kernel<<<tr_blocks(M * C), TR_T, 0, (cudaStream_t)stream>>>(p, d_grad_out, d_grad_grid);
    TR_CHECK_LAUNCH();
}

}  // extern "C"

// k4_ops.cu -- the op-level surface of the reference extension `render_utils_cuda`
// (lib/cuda/render_utils.cpp:170-184): the 13 functions the reference's training loop calls through
// autograd shims (lib/dvgo.py:453-511).  SURVEY.md section 8(f-2): needed for training compatibility,
// not for inference (the fused marchers never materialise these intermediates).
//
// Each kernel restates one reference kernel with its floating-point shape pinned by explicit
// intrinsics (nvcc contracts a*b+c in the reference build), so results are
// bit-identical to the reference extension (tests/test_gpu_ops_module.py compares against
// oracle/_ref).  Differences by design: everything is enqueued on the caller's stream, nothing
// synchronises (the reference has two hidden host syncs, render_utils_kernel.cu:212,635 -- the
// point count of sample_pts_on_rays is instead returned through a device scalar the Python wrapper
// reads, and the last segment's i_end is fixed up on the device).
#include "k4_internal.cuh"

namespace {

constexpr int OPS_T = 256;
inline unsigned ops_blocks(long long n) { return (unsigned)((n + OPS_T - 1) / OPS_T); }

// render_utils_kernel.cu:12-35
__global__ void op_infer_t_minmax(const float* __restrict__ ro, const float* __restrict__ rd,
                                  const float* __restrict__ mn, const float* __restrict__ mx,
                                  float near_, float far_, long long n, float* __restrict__ t_min, float* __restrict__ t_max) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float dx = rd[3 * i], dy = rd[3 * i + 1], dz = rd[3 * i + 2];
    const float vx = (dx == 0.f) ? 1e-6f : dx, vy = (dy == 0.f) ? 1e-6f : dy, vz = (dz == 0.f) ? 1e-6f : dz;
    const float ax = __fdiv_rn(__fsub_rn(mx[0], ro[3 * i]), vx), bx = __fdiv_rn(__fsub_rn(mn[0], ro[3 * i]), vx);
    const float ay = __fdiv_rn(__fsub_rn(mx[1], ro[3 * i + 1]), vy), by = __fdiv_rn(__fsub_rn(mn[1], ro[3 * i + 1]), vy);
    const float az = __fdiv_rn(__fsub_rn(mx[2], ro[3 * i + 2]), vz), bz = __fdiv_rn(__fsub_rn(mn[2], ro[3 * i + 2]), vz);
    t_min[i] = fmaxf(fminf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), far_), near_);
    t_max[i] = fmaxf(fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)), far_), near_);
}

__device__ __forceinline__ float op_rnorm(const float* d) {
    float s = __fmul_rn(d[1], d[1]);          // reference SASS: FMUL y,y ; FFMA x,x,. ; FFMA z,z,.
    s = __fmaf_rn(d[0], d[0], s);
    s = __fmaf_rn(d[2], d[2], s);
    return __fsqrt_rn(s);
}

// render_utils_kernel.cu:38-55
__global__ void op_infer_n_samples(const float* __restrict__ rd, const float* __restrict__ t_min, const float* __restrict__ t_max,
                                   float stepdist, long long n, long long* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float rnorm = op_rnorm(rd + 3 * i);
    const float c = ceilf(__fdiv_rn(__fmul_rn(__fsub_rn(t_max[i], t_min[i]), rnorm), stepdist));
    const double m = fmax((double)c, 1.0);
    out[i] = (long long)m;
}

// render_utils_kernel.cu:58-79
__global__ void op_infer_ray_start_dir(const float* __restrict__ ro, const float* __restrict__ rd, const float* __restrict__ t_min,
                                       long long n, float* __restrict__ start, float* __restrict__ dir) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float rnorm = op_rnorm(rd + 3 * i);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        start[3 * i + c] = __fmaf_rn(rd[3 * i + c], t_min[i], ro[3 * i + c]);
        dir[3 * i + c] = __fdiv_rn(rd[3 * i + c], rnorm);
    }
}

// render_utils_kernel.cu:144-164 (scatter-1 + cumsum + __set_step_id) as one lookup per point
__global__ void op_fill_ids(const long long* __restrict__ cumsum, long long n_rays, long long total,
                            long long* __restrict__ ray_id, long long* __restrict__ step_id) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    long long lo = 0, hi = n_rays - 1;          // first ray whose inclusive cumsum exceeds i
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (cumsum[mid] > i) hi = mid; else lo = mid + 1;
    }
    ray_id[i] = lo;
    step_id[i] = i - (lo ? cumsum[lo - 1] : 0);
}

// render_utils_kernel.cu:167-194
__global__ void op_sample_pts(const float* __restrict__ start, const float* __restrict__ dir,
                              const float* __restrict__ mn, const float* __restrict__ mx,
                              const long long* __restrict__ ray_id, const long long* __restrict__ step_id,
                              float stepdist, long long total, float* __restrict__ pts, bool* __restrict__ mask) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int r = (int)ray_id[i], st = (int)step_id[i];
    const float dist = __fmul_rn(stepdist, (float)st);
    const float px = __fmaf_rn(dir[3 * r], dist, start[3 * r]);
    const float py = __fmaf_rn(dir[3 * r + 1], dist, start[3 * r + 1]);
    const float pz = __fmaf_rn(dir[3 * r + 2], dist, start[3 * r + 2]);
    pts[3 * i] = px; pts[3 * i + 1] = py; pts[3 * i + 2] = pz;
    mask[i] = (mn[0] > px) | (mn[1] > py) | (mn[2] > pz) | (mx[0] < px) | (mx[1] < py) | (mx[2] < pz);
}

// render_utils_kernel.cu:245-270
__global__ void op_sample_ndc(const float* __restrict__ ro, const float* __restrict__ rd,
                              const float* __restrict__ mn, const float* __restrict__ mx,
                              int N_samples, long long n_rays, float* __restrict__ pts, bool* __restrict__ mask) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N_samples * n_rays) return;
    const long long r = i / N_samples;
    const int st = (int)(i - r * N_samples);
    const float dist = __fdiv_rn((float)st, (float)(N_samples - 1));
    const float px = __fmaf_rn(rd[3 * r], dist, ro[3 * r]);
    const float py = __fmaf_rn(rd[3 * r + 1], dist, ro[3 * r + 1]);
    const float pz = __fmaf_rn(rd[3 * r + 2], dist, ro[3 * r + 2]);
    pts[3 * i] = px; pts[3 * i + 1] = py; pts[3 * i + 2] = pz;
    mask[i] = (mn[0] > px) | (mn[1] > py) | (mn[2] > pz) | (mx[0] < px) | (mx[1] < py) | (mx[2] < pz);
}

// render_utils_kernel.cu:301-340 (inverse-sphere background samples; the double literals make most of
// the arithmetic double precision in the reference)
__global__ void op_sample_bg(const float* __restrict__ ro, const float* __restrict__ rd, const float* __restrict__ t_max,
                             float bg_preserve, int N_samples, long long n_rays, float* __restrict__ pts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N_samples * n_rays) return;
    const long long r = i / N_samples;
    const int st = (int)(i - r * N_samples);
    const float t_inner = t_max[r];
    const float q = __fdiv_rn((float)st, (float)N_samples);
    const float ori_t_outer = (float)(((double)t_inner - 1.0) + 1.0 / (1.0 - (double)q));
    const float ox = __fmaf_rn(rd[3 * r], ori_t_outer, ro[3 * r]);
    const float oy = __fmaf_rn(rd[3 * r + 1], ori_t_outer, ro[3 * r + 1]);
    const float oz = __fmaf_rn(rd[3 * r + 2], ori_t_outer, ro[3 * r + 2]);
    float nn = __fmul_rn(oy, oy);               // norm3: x*x + y*y + z*z, contracted like the other norms
    nn = __fmaf_rn(ox, ox, nn);
    nn = __fmaf_rn(oz, oz, nn);
    const float t_outer = __fsqrt_rn(nn);
    const float om = fmaxf(fabsf(ox), fmaxf(fabsf(oy), fabsf(oz)));
    const float R = __fdiv_rn(t_outer, om);
    // R*R / (t*t) * (1.-bg) + R / t * bg  -- float products/quotients, double scaling and sum
    const double q2 = (double)__fdiv_rn(__fmul_rn(R, R), __fmul_rn(t_outer, t_outer));
    const double b = (double)__fmul_rn(__fdiv_rn(R, t_outer), bg_preserve);
    const float o2i = (float)fma(q2, 1.0 - (double)bg_preserve, b);          // the double a*b+c is contracted too
    pts[3 * i] = __fmul_rn(ox, o2i); pts[3 * i + 1] = __fmul_rn(oy, o2i); pts[3 * i + 2] = __fmul_rn(oz, o2i);
}

// render_utils_kernel.cu:373-392 (out is pre-zeroed by the caller)
__global__ void op_maskcache(const bool* __restrict__ world, const float* __restrict__ xyz, bool* __restrict__ out,
                             const float* __restrict__ scale, const float* __restrict__ shift,
                             int si, int sj, int sk, long long n) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int i = (int)roundf(__fmaf_rn(xyz[3 * p], scale[0], shift[0]));
    const int j = (int)roundf(__fmaf_rn(xyz[3 * p + 1], scale[1], shift[1]));
    const int k = (int)roundf(__fmaf_rn(xyz[3 * p + 2], scale[2], shift[2]));
    if ((0 <= i) & (i < si) & (0 <= j) & (j < sj) & (0 <= k) & (k < sk)) out[p] = world[((size_t)i * sj + j) * sk + k];
}

// render_utils_kernel.cu:431-458
__global__ void op_raw2alpha(const float* __restrict__ density, float shift, float interval, const float* __restrict__ interval_v,
                             long long n, float* __restrict__ exp_d, float* __restrict__ alpha) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float e = expf(__fadd_rn(density[i], shift));
    exp_d[i] = e;
    const float iv = interval_v ? interval_v[i] : interval;
    alpha[i] = __fsub_rn(1.f, powf(__fadd_rn(1.f, e), -iv));
}

// render_utils_kernel.cu:507-530:  min(e, 1e10) * pow(1+e, -interval-1) * interval * g   (double product chain)
__global__ void op_raw2alpha_bwd(const float* __restrict__ exp_d, const float* __restrict__ g, float interval,
                                 const float* __restrict__ interval_v, long long n, float* __restrict__ grad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float iv = interval_v ? interval_v[i] : interval;
    const float e = exp_d[i];
    const double m = fmin((double)e, 1e10);
    const float pw = powf(__fadd_rn(1.f, e), __fsub_rn(-iv, 1.f));
    grad[i] = (float)(((m * (double)pw) * (double)iv) * (double)g[i]);
}

// render_utils_kernel.cu:607-617 + the host-side fix-up of :635 done on the device
__global__ void op_segments(const long long* __restrict__ ray_id, long long n_pts, long long* __restrict__ i_start, long long* __restrict__ i_end) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pts) return;
    if (i > 0 && ray_id[i] != ray_id[i - 1]) { i_start[ray_id[i]] = i; i_end[ray_id[i - 1]] = i; }
    if (i == n_pts - 1) i_end[ray_id[i]] = n_pts;
}

// render_utils_kernel.cu:577-605
__global__ void op_alpha2weight(const float* __restrict__ alpha, long long n_rays, float* __restrict__ weight, float* __restrict__ T,
                                float* __restrict__ last, const long long* __restrict__ i_start, long long* __restrict__ i_end) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const long long i_s = i_start[r], i_e = i_end[r];
    float Tc = 1.f;
    long long i;
    for (i = i_s; i < i_e; ++i) {
        T[i] = Tc;
        weight[i] = __fmul_rn(Tc, alpha[i]);
        Tc = (float)((double)Tc * (1.0 - (double)alpha[i]));
        if ((double)Tc < 1e-3) { i += 1; break; }
    }
    i_end[r] = i;
    last[r] = Tc;
}

// render_utils_kernel.cu:654-677
__global__ void op_alpha2weight_bwd(const float* __restrict__ alpha, const float* __restrict__ weight, const float* __restrict__ T,
                                    const float* __restrict__ last, const long long* __restrict__ i_start, const long long* __restrict__ i_end,
                                    long long n_rays, const float* __restrict__ gw, const float* __restrict__ gl, float* __restrict__ grad) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const long long i_s = i_start[r], i_e = i_end[r];
    float back = __fmul_rn(gl[r], last[r]);
    for (long long i = i_e - 1; i >= i_s; --i) {
        const double den = (double)__fsub_rn(1.f, alpha[i]) + 1e-10;
        grad[i] = (float)((double)__fmul_rn(gw[i], T[i]) - (double)back / den);
        back = __fmaf_rn(gw[i], weight[i], back);
    }
}

}  // namespace

#define OPS_LAUNCH(kern, n, ...)                                                           \
    do {                                                                                   \
        if ((n) > 0) {                                                                     \
            kern<<<ops_blocks(n), OPS_T, 0, (cudaStream_t)stream>>>(__VA_ARGS__);          \
            K4_CUDA_TRY(cudaGetLastError());                                               \
        }                                                                                  \
        return K4_OK;                                                                      \
    } while (0)

extern "C" {

int k4_op_infer_t_minmax(const float* ro, const float* rd, const float* d_min, const float* d_max, float near_, float far_,
                         int64_t n, float* t_min, float* t_max, k4_stream_t stream) {
    OPS_LAUNCH(op_infer_t_minmax, n, ro, rd, d_min, d_max, near_, far_, (long long)n, t_min, t_max);
}
int k4_op_infer_n_samples(const float* rd, const float* t_min, const float* t_max, float stepdist, int64_t n, int64_t* out, k4_stream_t stream) {
    OPS_LAUNCH(op_infer_n_samples, n, rd, t_min, t_max, stepdist, (long long)n, (long long*)out);
}
int k4_op_infer_ray_start_dir(const float* ro, const float* rd, const float* t_min, int64_t n, float* start, float* dir, k4_stream_t stream) {
    OPS_LAUNCH(op_infer_ray_start_dir, n, ro, rd, t_min, (long long)n, start, dir);
}
int k4_op_fill_ray_step_ids(const int64_t* cumsum, int64_t n_rays, int64_t total, int64_t* ray_id, int64_t* step_id, k4_stream_t stream) {
    OPS_LAUNCH(op_fill_ids, total, (const long long*)cumsum, (long long)n_rays, (long long)total, (long long*)ray_id, (long long*)step_id);
}
int k4_op_sample_pts(const float* start, const float* dir, const float* d_min, const float* d_max, const int64_t* ray_id,
                     const int64_t* step_id, float stepdist, int64_t total, float* pts, uint8_t* mask, k4_stream_t stream) {
    OPS_LAUNCH(op_sample_pts, total, start, dir, d_min, d_max, (const long long*)ray_id, (const long long*)step_id, stepdist,
               (long long)total, pts, (bool*)mask);
}
int k4_op_sample_ndc_pts(const float* ro, const float* rd, const float* d_min, const float* d_max, int32_t N_samples, int64_t n_rays,
                         float* pts, uint8_t* mask, k4_stream_t stream) {
    OPS_LAUNCH(op_sample_ndc, (long long)N_samples * n_rays, ro, rd, d_min, d_max, N_samples, (long long)n_rays, pts, (bool*)mask);
}
int k4_op_sample_bg_pts(const float* ro, const float* rd, const float* t_max, float bg_preserve, int32_t N_samples, int64_t n_rays,
                        float* pts, k4_stream_t stream) {
    OPS_LAUNCH(op_sample_bg, (long long)N_samples * n_rays, ro, rd, t_max, bg_preserve, N_samples, (long long)n_rays, pts);
}
int k4_op_maskcache_lookup(const uint8_t* world, const float* xyz, uint8_t* out, const float* scale, const float* shift,
                           int32_t si, int32_t sj, int32_t sk, int64_t n, k4_stream_t stream) {
    OPS_LAUNCH(op_maskcache, n, (const bool*)world, xyz, (bool*)out, scale, shift, si, sj, sk, (long long)n);
}
int k4_op_raw2alpha(const float* density, float shift, float interval, const float* interval_v, int64_t n, float* exp_d, float* alpha,
                    k4_stream_t stream) {
    OPS_LAUNCH(op_raw2alpha, n, density, shift, interval, interval_v, (long long)n, exp_d, alpha);
}
int k4_op_raw2alpha_backward(const float* exp_d, const float* grad_back, float interval, const float* interval_v, int64_t n, float* grad,
                             k4_stream_t stream) {
    OPS_LAUNCH(op_raw2alpha_bwd, n, exp_d, grad_back, interval, interval_v, (long long)n, grad);
}
int k4_op_alpha2weight(const float* alpha, const int64_t* ray_id, int64_t n_rays, int64_t n_pts, float* weight, float* T, float* last,
                       int64_t* i_start, int64_t* i_end, k4_stream_t stream) {
    if (n_pts <= 0) return K4_OK;
    op_segments<<<ops_blocks(n_pts), OPS_T, 0, (cudaStream_t)stream>>>((const long long*)ray_id, (long long)n_pts, (long long*)i_start, (long long*)i_end);
    K4_CUDA_TRY(cudaGetLastError());
    OPS_LAUNCH(op_alpha2weight, n_rays, alpha, (long long)n_rays, weight, T, last, (const long long*)i_start, (long long*)i_end);
}
int k4_op_alpha2weight_backward(const float* alpha, const float* weight, const float* T, const float* last, const int64_t* i_start,
                                const int64_t* i_end, int64_t n_rays, const float* grad_weights, const float* grad_last, float* grad,
                                k4_stream_t stream) {
    OPS_LAUNCH(op_alpha2weight_bwd, n_rays, alpha, weight, T, last, (const long long*)i_start, (const long long*)i_end, (long long)n_rays,
               grad_weights, grad_last, grad);
}

}  // extern "C"

/*
 * oracle/render_utils_ref.c -- TEST INFRASTRUCTURE ONLY (the parity oracle), never the product path.
 *
 * Plain-C, CPU restatement of the forward ray-march ops of the reference extension
 * /root/reference/lib/cuda/render_utils_kernel.cu (bound in lib/cuda/render_utils.cpp:170-184).
 * Each function cites the reference kernel it follows.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Floating-point shape.  The reference is compiled by nvcc with its defaults (-fmad=true, no
 * fast-math), so every `a*b + c` in one expression is contracted into a single-rounding FMA on
 * the device.  The geometric results (t_min/t_max, step counts, in-box masks, occupancy indices)
 * are compared BIT-EXACTLY, so the contraction is restated explicitly with fmaf() at the sites
 * where nvcc fuses (verified against the SASS of the reference build) and
 * this file must be compiled with -ffp-contract=off so that gcc adds none of its own.
 * expf/powf come from the host libm here and from the CUDA math library on the device: those two
 * differ by <= 2 ulp, which is why alpha/weights are compared within a tolerance, not bit-exactly.
 *
 * Pinning status: the reference ships no tests / golden vectors.  This
 * restatement is pinned on the GPU box against oracle/_ref/render_utils_cuda.so, i.e. the
 * reference's own kernels compiled from /root/reference (tests/test_gpu_ref_ops.py).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define K4O_API __attribute__((visibility("default")))

K4O_API int k4o_abi_version(void) { return 1; }

K4O_API int k4o_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

K4O_API void k4o_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* render_utils_kernel.cu:12-35 infer_t_minmax_cuda_kernel.
 * `(d==0) ? 1e-6 : d` is a double expression narrowed to float; min/max are fminf/fmaxf. */
K4O_API void k4o_infer_t_minmax(const float* rays_o, const float* rays_d,
                                const float* xyz_min, const float* xyz_max,
                                float near, float far, int64_t n_rays,
                                float* t_min, float* t_max) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_rays; ++r) {
        const float* o = rays_o + 3 * r;
        const float* d = rays_d + 3 * r;
        float vx = (d[0] == 0.f) ? (float)1e-6 : d[0];
        float vy = (d[1] == 0.f) ? (float)1e-6 : d[1];
        float vz = (d[2] == 0.f) ? (float)1e-6 : d[2];
        float ax = (xyz_max[0] - o[0]) / vx;
        float ay = (xyz_max[1] - o[1]) / vy;
        float az = (xyz_max[2] - o[2]) / vz;
        float bx = (xyz_min[0] - o[0]) / vx;
        float by = (xyz_min[1] - o[1]) / vy;
        float bz = (xyz_min[2] - o[2]) / vz;
        t_min[r] = fmaxf(fminf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), far), near);
        t_max[r] = fmaxf(fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)), far), near);
    }
}

/* squared norm as nvcc contracts `x*x + y*y + z*z` in the reference build (SASS of
 * infer_n_samples / infer_ray_start_dir: FMUL y,y ; FFMA x,x,. ; FFMA z,z,.): fma(z,z, fma(x,x, y*y)). */
static inline float k4o_rnorm(const float* d) {
    float s = d[1] * d[1];
    s = fmaf(d[0], d[0], s);
    s = fmaf(d[2], d[2], s);
    return sqrtf(s);
}

/* render_utils_kernel.cu:38-55 infer_n_samples_cuda_kernel.
 * n = max(ceil((t_max-t_min)*rnorm/stepdist), 1.) stored as int64. */
K4O_API void k4o_infer_n_samples(const float* rays_d, const float* t_min, const float* t_max,
                                 float stepdist, int64_t n_rays, int64_t* n_samples) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_rays; ++r) {
        float rnorm = k4o_rnorm(rays_d + 3 * r);
        float x = ceilf((t_max[r] - t_min[r]) * rnorm / stepdist);
        double m = ((double)x > 1.0) ? (double)x : 1.0;   /* max(float, double) -> double */
        n_samples[r] = (int64_t)m;
    }
}

/* render_utils_kernel.cu:58-79 infer_ray_start_dir_cuda_kernel.
 * start = o + d*t_min (one FMA per component), dir = d / rnorm. */
K4O_API void k4o_infer_ray_start_dir(const float* rays_o, const float* rays_d, const float* t_min,
                                     int64_t n_rays, float* rays_start, float* rays_dir) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_rays; ++r) {
        const float* o = rays_o + 3 * r;
        const float* d = rays_d + 3 * r;
        float rnorm = k4o_rnorm(d);
        for (int c = 0; c < 3; ++c) {
            rays_start[3 * r + c] = fmaf(d[c], t_min[r], o[c]);
            rays_dir[3 * r + c] = d[c] / rnorm;
        }
    }
}

/* render_utils_kernel.cu:144-164 (__set_1_at_ray_seg_start + cumsum + __set_step_id):
 * flat ray_id / step_id lists from per-ray step counts. */
K4O_API void k4o_fill_ray_step_ids(const int64_t* n_steps, int64_t n_rays,
                                   int64_t* ray_id, int64_t* step_id) {
    int64_t k = 0;
    for (int64_t r = 0; r < n_rays; ++r)
        for (int64_t s = 0; s < n_steps[r]; ++s) { ray_id[k] = r; step_id[k] = s; ++k; }
}

/* render_utils_kernel.cu:167-194 sample_pts_on_rays_cuda_kernel.
 * dist = stepdist*(float)i_step; p = start + dir*dist (FMA); mask = any(min>p | max<p). */
K4O_API void k4o_sample_pts_on_rays(const float* rays_start, const float* rays_dir,
                                    const float* xyz_min, const float* xyz_max,
                                    const int64_t* ray_id, const int64_t* step_id,
                                    float stepdist, int64_t total_len,
                                    float* rays_pts, uint8_t* mask_outbbox) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < total_len; ++i) {
        const int i_ray = (int)ray_id[i];
        const int i_step = (int)step_id[i];
        const float dist = stepdist * (float)i_step;
        const float px = fmaf(rays_dir[3 * i_ray + 0], dist, rays_start[3 * i_ray + 0]);
        const float py = fmaf(rays_dir[3 * i_ray + 1], dist, rays_start[3 * i_ray + 1]);
        const float pz = fmaf(rays_dir[3 * i_ray + 2], dist, rays_start[3 * i_ray + 2]);
        rays_pts[3 * i + 0] = px;
        rays_pts[3 * i + 1] = py;
        rays_pts[3 * i + 2] = pz;
        mask_outbbox[i] = (uint8_t)((xyz_min[0] > px) | (xyz_min[1] > py) | (xyz_min[2] > pz) |
                                    (xyz_max[0] < px) | (xyz_max[1] < py) | (xyz_max[2] < pz));
    }
}

/* render_utils_kernel.cu:245-270 sample_ndc_pts_on_rays_cuda_kernel.
 * dist = (float)i_step / (N_samples-1); p = o + d*dist (FMA); dense [n_rays, N_samples]. */
K4O_API void k4o_sample_ndc_pts_on_rays(const float* rays_o, const float* rays_d,
                                        const float* xyz_min, const float* xyz_max,
                                        int N_samples, int64_t n_rays,
                                        float* rays_pts, uint8_t* mask_outbbox) {
    const int64_t total = (int64_t)N_samples * n_rays;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < total; ++i) {
        const int64_t i_ray = i / N_samples;
        const int i_step = (int)(i % N_samples);
        const float dist = ((float)i_step) / (float)(N_samples - 1);
        const float px = fmaf(rays_d[3 * i_ray + 0], dist, rays_o[3 * i_ray + 0]);
        const float py = fmaf(rays_d[3 * i_ray + 1], dist, rays_o[3 * i_ray + 1]);
        const float pz = fmaf(rays_d[3 * i_ray + 2], dist, rays_o[3 * i_ray + 2]);
        rays_pts[3 * i + 0] = px;
        rays_pts[3 * i + 1] = py;
        rays_pts[3 * i + 2] = pz;
        mask_outbbox[i] = (uint8_t)((xyz_min[0] > px) | (xyz_min[1] > py) | (xyz_min[2] > pz) |
                                    (xyz_max[0] < px) | (xyz_max[1] < py) | (xyz_max[2] < pz));
    }
}

/* float -> int as the device does it (cvt.rzi.s32.f32 saturates, NaN -> 0). */
static inline int k4o_f2i(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (-2147483647 - 1);
    return (int)v;
}

/* render_utils_kernel.cu:373-392 maskcache_lookup_cuda_kernel.
 * ijk = round(p*scale + shift) with the FMA contracted, C round() = half away from zero;
 * out is pre-zeroed by the host wrapper (:405), written only when ijk is inside. */
K4O_API void k4o_maskcache_lookup(const uint8_t* world, const float* xyz, uint8_t* out,
                                  const float* scale, const float* shift,
                                  int sz_i, int sz_j, int sz_k, int64_t n_pts) {
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n_pts; ++p) {
        const int i = k4o_f2i(roundf(fmaf(xyz[3 * p + 0], scale[0], shift[0])));
        const int j = k4o_f2i(roundf(fmaf(xyz[3 * p + 1], scale[1], shift[1])));
        const int k = k4o_f2i(roundf(fmaf(xyz[3 * p + 2], scale[2], shift[2])));
        uint8_t v = 0;
        if (0 <= i && i < sz_i && 0 <= j && j < sz_j && 0 <= k && k < sz_k)
            v = world[(int64_t)i * sz_j * sz_k + (int64_t)j * sz_k + k];
        out[p] = v;
    }
}

/* render_utils_kernel.cu:431-443 raw2alpha_cuda_kernel.
 * e = exp(d + shift) (may be inf); alpha = 1 - pow(1+e, -interval). */
K4O_API void k4o_raw2alpha(const float* density, float shift, float interval, int64_t n_pts,
                           float* exp_d, float* alpha) {
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n_pts; ++p) {
        const float e = expf(density[p] + shift);
        exp_d[p] = e;
        alpha[p] = 1.f - powf(1.f + e, -interval);
    }
}

/* render_utils_kernel.cu:577-651 alpha2weight_cuda (+ __set_i_for_segment_start_end :607-617,
 * the host-side i_end fix-up :635, and the serial per-ray loop :586-604).
 * weight/T/alphainv_last/i_start/i_end must be pre-filled by the caller with 0/1/1/0/0 exactly
 * as the reference wrapper does (:624-628).  T_cum is float but the factor (1. - alpha) is a
 * double expression: T_cum = (float)((double)T_cum * (1.0 - (double)alpha)); the early-out test
 * compares the float against the double literal 1e-3. */
K4O_API void k4o_alpha2weight(const float* alpha, const int64_t* ray_id, int64_t n_rays, int64_t n_pts,
                              float* weight, float* T, float* alphainv_last,
                              int64_t* i_start, int64_t* i_end) {
    if (n_pts == 0) return;
    for (int64_t i = 1; i < n_pts; ++i) {
        if (ray_id[i] != ray_id[i - 1]) {
            i_start[ray_id[i]] = i;
            i_end[ray_id[i - 1]] = i;
        }
    }
    i_end[ray_id[n_pts - 1]] = n_pts;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_rays; ++r) {
        const int64_t i_s = i_start[r];
        const int64_t i_e_max = i_end[r];
        float T_cum = 1.f;
        int64_t i;
        for (i = i_s; i < i_e_max; ++i) {
            T[i] = T_cum;
            weight[i] = T_cum * alpha[i];
            T_cum = (float)((double)T_cum * (1.0 - (double)alpha[i]));
            if ((double)T_cum < 1e-3) { i += 1; break; }
        }
        i_end[r] = i;
        alphainv_last[r] = T_cum;
    }
}

/* lib/cuda/ub360_utils_kernel.cu:12-32 cumdist_thres_cuda_kernel: per ray, accumulate the distances
 * between consecutive samples; emit (and reset) whenever the sum exceeds thres. */
K4O_API void k4o_cumdist_thres(const float* dist, float thres, int64_t n_rays, int64_t n_pts, uint8_t* mask) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_rays; ++r) {
        float cum = 0.f;
        for (int64_t i = r * n_pts; i < (r + 1) * n_pts; ++i) {
            cum += dist[i];
            const int over = cum > thres;
            cum *= (float)(!over);
            mask[i] = (uint8_t)over;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Training-side element-wise kernels
 * ------------------------------------------------------------------------------------------ */
static inline float k4o_clamp1(float v) { return fminf(fmaxf(v, -1.f), 1.f); }

/* lib/cuda/total_variation_kernel.cu:13-66.  Reference build: every term is a separate product added to a
 * running sum that starts at +0 (read from its SASS), then grad += sum. */
K4O_API void k4o_total_variation_add_grad(const float* param, float* grad, float wx, float wy, float wz, int dense_mode,
                                          int64_t n, int64_t szi, int64_t szj, int64_t szk) {
    wx /= 6; wy /= 6; wz /= 6;
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < n; ++idx) {
        const float g0 = grad[idx];
        if (!dense_mode && g0 == 0.f) continue;
        const int64_t k = idx % szk, j = idx / szk % szj, i = idx / szk / szj % szi;
        const float p = param[idx];
        float acc = 0.f;
        acc = acc + ((k == 0) ? 0.f : wx * k4o_clamp1(p - param[idx - 1]));
        acc = acc + ((k == szk - 1) ? 0.f : wx * k4o_clamp1(p - param[idx + 1]));
        acc = acc + ((j == 0) ? 0.f : wy * k4o_clamp1(p - param[idx - szk]));
        acc = acc + ((j == szj - 1) ? 0.f : wy * k4o_clamp1(p - param[idx + szk]));
        acc = acc + ((i == 0) ? 0.f : wz * k4o_clamp1(p - param[idx - szk * szj]));
        acc = acc + ((i == szi - 1) ? 0.f : wz * k4o_clamp1(p - param[idx + szk * szj]));
        grad[idx] = g0 + acc;
    }
}

/* lib/cuda/adam_upd_kernel.cu:8-136 (adam_upd / masked_adam_upd / adam_upd_with_perlr).  Reference build:
 * m = fma(m, b1, (1-b1)*g); v = fma(v, b2, ((1-b2)*g)*g); p -= (step_size [* perlr]) * m / (sqrt(v) + eps),
 * IEEE sqrt and division; step_size in fp32 on the host (:76). */
K4O_API void k4o_adam_upd(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const float* perlr, int64_t n,
                          int step, float beta1, float beta2, float lr, float eps, int skip_zero_grad) {
    const float step_size = lr * sqrtf(1 - powf(beta2, (float)step)) / (1 - powf(beta1, (float)step));
    const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const float g = grad[i];
        if (!perlr && skip_zero_grad && g == 0.f) continue;
        const float m = fmaf(beta1, exp_avg[i], omb1 * g);
        const float v = fmaf(beta2, exp_avg_sq[i], (omb2 * g) * g);
        const float num = perlr ? (step_size * perlr[i]) * m : step_size * m;
        exp_avg[i] = m;
        exp_avg_sq[i] = v;
        param[i] = param[i] - num / (sqrtf(v) + eps);
    }
}

/* ------------------------------------------------------------------------------------------
 * Counting helper used for the roofline figure: S_m, S_d, S_c are the
 * number of samples reaching the mask lookup / density fetch / feature fetch.  Pure bookkeeping
 * on masks produced by the pipeline; no reference counterpart.
 * ------------------------------------------------------------------------------------------ */
K4O_API int64_t k4o_count_true(const uint8_t* m, int64_t n) {
    int64_t c = 0;
#pragma omp parallel for reduction(+ : c) schedule(static)
    for (int64_t i = 0; i < n; ++i) c += (m[i] != 0);
    return c;
}

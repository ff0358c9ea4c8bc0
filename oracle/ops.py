"""Op-level oracle backends with the reference extension's interface -- TEST INFRASTRUCTURE ONLY.

Both classes expose the forward functions of ``render_utils_cuda`` (reference binding table:
lib/cuda/render_utils.cpp:170-184) with the same names, argument order and return lists:

``CpuOps``     torch-CPU tensors in/out, arithmetic in oracle/render_utils_ref.c (ctypes).
``RefExtOps``  CUDA tensors in/out, arithmetic = the reference's OWN kernels, compiled from
               /root/reference by oracle/build_ref.py into oracle/_ref/render_utils_cuda.so.
               Only usable on a GPU box; this is what pins CpuOps (tests/test_gpu_ref_ops.py).
"""
import ctypes
import importlib.util
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c_oracle(force=False):
    """Compile oracle/render_utils_ref.c -> oracle/libk4oracle.so (gcc, seconds)."""
    so = os.path.join(HERE, 'libk4oracle.so')
    src = os.path.join(HERE, 'render_utils_ref.c')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-C', HERE, 'libk4oracle.so'])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c_oracle())
        _LIB.k4o_count_true.restype = ctypes.c_int64
        _LIB.k4o_num_threads.restype = ctypes.c_int
    return _LIB


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _f32(t):
    t = t.detach()
    assert t.device.type == 'cpu'
    return t.to(torch.float32).contiguous()


def set_num_threads(n):
    _lib().k4o_set_num_threads(ctypes.c_int(int(n)))
    torch.set_num_threads(int(n))


def num_threads():
    return int(_lib().k4o_num_threads())


class CpuOps:
    """CPU restatement of render_utils_cuda's forward functions (see module docstring)."""
    device = 'cpu'

    @staticmethod
    def infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far):
        rays_o, rays_d, xyz_min, xyz_max = map(_f32, (rays_o, rays_d, xyz_min, xyz_max))
        n = rays_o.shape[0]
        t_min = torch.empty(n)
        t_max = torch.empty(n)
        _lib().k4o_infer_t_minmax(_p(rays_o), _p(rays_d), _p(xyz_min), _p(xyz_max),
                                  ctypes.c_float(near), ctypes.c_float(far), ctypes.c_int64(n),
                                  _p(t_min), _p(t_max))
        return [t_min, t_max]

    @staticmethod
    def infer_n_samples(rays_d, t_min, t_max, stepdist):
        rays_d, t_min, t_max = map(_f32, (rays_d, t_min, t_max))
        n = t_min.shape[0]
        out = torch.empty(n, dtype=torch.int64)
        _lib().k4o_infer_n_samples(_p(rays_d), _p(t_min), _p(t_max), ctypes.c_float(float(stepdist)),
                                   ctypes.c_int64(n), _p(out))
        return out

    @staticmethod
    def infer_ray_start_dir(rays_o, rays_d, t_min):
        rays_o, rays_d, t_min = map(_f32, (rays_o, rays_d, t_min))
        n = rays_o.shape[0]
        start = torch.empty_like(rays_o)
        rdir = torch.empty_like(rays_o)
        _lib().k4o_infer_ray_start_dir(_p(rays_o), _p(rays_d), _p(t_min), ctypes.c_int64(n),
                                       _p(start), _p(rdir))
        return [start, rdir]

    @staticmethod
    def sample_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist):
        """render_utils_kernel.cu:196-242 sample_pts_on_rays_cuda (host wrapper)."""
        rays_o, rays_d, xyz_min, xyz_max = map(_f32, (rays_o, rays_d, xyz_min, xyz_max))
        stepdist = float(stepdist)
        t_min, t_max = CpuOps.infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far)
        n_steps = CpuOps.infer_n_samples(rays_d, t_min, t_max, stepdist)
        total = int(n_steps.sum().item())
        ray_id = torch.empty(total, dtype=torch.int64)
        step_id = torch.empty(total, dtype=torch.int64)
        _lib().k4o_fill_ray_step_ids(_p(n_steps), ctypes.c_int64(rays_o.shape[0]), _p(ray_id), _p(step_id))
        start, rdir = CpuOps.infer_ray_start_dir(rays_o, rays_d, t_min)
        pts = torch.empty(total, 3)
        mask = torch.empty(total, dtype=torch.bool)
        _lib().k4o_sample_pts_on_rays(_p(start), _p(rdir), _p(xyz_min), _p(xyz_max), _p(ray_id), _p(step_id),
                                      ctypes.c_float(stepdist), ctypes.c_int64(total), _p(pts), _p(mask))
        return [pts, mask, ray_id, step_id, n_steps, t_min, t_max]

    @staticmethod
    def sample_ndc_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, N_samples):
        rays_o, rays_d, xyz_min, xyz_max = map(_f32, (rays_o, rays_d, xyz_min, xyz_max))
        n = rays_o.shape[0]
        pts = torch.empty(n, N_samples, 3)
        mask = torch.empty(n, N_samples, dtype=torch.bool)
        _lib().k4o_sample_ndc_pts_on_rays(_p(rays_o), _p(rays_d), _p(xyz_min), _p(xyz_max),
                                          ctypes.c_int(int(N_samples)), ctypes.c_int64(n), _p(pts), _p(mask))
        return [pts, mask]

    @staticmethod
    def maskcache_lookup(world, xyz, xyz2ijk_scale, xyz2ijk_shift):
        world = world.detach().to(torch.bool).contiguous()
        xyz, scale, shift = map(_f32, (xyz, xyz2ijk_scale, xyz2ijk_shift))
        n = xyz.shape[0]
        out = torch.zeros(n, dtype=torch.bool)
        if n:
            _lib().k4o_maskcache_lookup(_p(world), _p(xyz), _p(out), _p(scale), _p(shift),
                                        ctypes.c_int(world.shape[0]), ctypes.c_int(world.shape[1]),
                                        ctypes.c_int(world.shape[2]), ctypes.c_int64(n))
        return out

    @staticmethod
    def raw2alpha(density, shift, interval):
        density = _f32(density)
        n = density.shape[0]
        exp_d = torch.empty_like(density)
        alpha = torch.empty_like(density)
        if n:
            _lib().k4o_raw2alpha(_p(density), ctypes.c_float(float(shift)), ctypes.c_float(float(interval)),
                                 ctypes.c_int64(n), _p(exp_d), _p(alpha))
        return [exp_d, alpha]

    @staticmethod
    def alpha2weight(alpha, ray_id, n_rays):
        alpha = _f32(alpha)
        ray_id = ray_id.detach().to(torch.int64).contiguous()
        n = alpha.shape[0]
        weight = torch.zeros_like(alpha)
        T = torch.ones_like(alpha)
        last = torch.ones(n_rays)
        i_start = torch.zeros(n_rays, dtype=torch.int64)
        i_end = torch.zeros(n_rays, dtype=torch.int64)
        if n:
            _lib().k4o_alpha2weight(_p(alpha), _p(ray_id), ctypes.c_int64(n_rays), ctypes.c_int64(n),
                                    _p(weight), _p(T), _p(last), _p(i_start), _p(i_end))
        return [weight, T, last, i_start, i_end]


    @staticmethod
    def cumdist_thres(dist, thres):
        """ub360_utils_cuda.cumdist_thres (lib/cuda/ub360_utils.cpp:20-22)."""
        dist = _f32(dist)
        mask = torch.zeros(dist.shape, dtype=torch.bool)
        if dist.numel():
            _lib().k4o_cumdist_thres(_p(dist), ctypes.c_float(float(thres)), ctypes.c_int64(dist.shape[0]),
                                     ctypes.c_int64(dist.shape[1]), _p(mask))
        return mask


    # ---- training-side element-wise ops (in place on CPU tensors), SURVEY.md section 8 f-4 ----
    @staticmethod
    def total_variation_add_grad(param, grad, wx, wy, wz, dense_mode):
        """total_variation_cuda.total_variation_add_grad (lib/cuda/total_variation.cpp:16-20)."""
        assert param.dtype == torch.float32 and grad.dtype == torch.float32 and param.is_contiguous() and grad.is_contiguous()
        _lib().k4o_total_variation_add_grad(_p(param), _p(grad), ctypes.c_float(wx), ctypes.c_float(wy), ctypes.c_float(wz),
                                            ctypes.c_int(int(bool(dense_mode))), ctypes.c_int64(param.numel()),
                                            ctypes.c_int64(param.shape[2]), ctypes.c_int64(param.shape[3]), ctypes.c_int64(param.shape[4]))

    @staticmethod
    def _adam(param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps, skip):
        for t in (param, grad, exp_avg, exp_avg_sq):
            assert t.dtype == torch.float32 and t.is_contiguous()
        _lib().k4o_adam_upd(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), _p(perlr) if perlr is not None else None,
                            ctypes.c_int64(param.numel()), ctypes.c_int(int(step)), ctypes.c_float(beta1), ctypes.c_float(beta2),
                            ctypes.c_float(lr), ctypes.c_float(eps), ctypes.c_int(int(skip)))

    @staticmethod
    def adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
        CpuOps._adam(param, grad, exp_avg, exp_avg_sq, None, step, beta1, beta2, lr, eps, 0)

    @staticmethod
    def masked_adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
        CpuOps._adam(param, grad, exp_avg, exp_avg_sq, None, step, beta1, beta2, lr, eps, 1)

    @staticmethod
    def adam_upd_with_perlr(param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps):
        CpuOps._adam(param, grad, exp_avg, exp_avg_sq, perlr.contiguous(), step, beta1, beta2, lr, eps, 0)


def ref_ext_path():
    return os.path.join(HERE, '_ref', 'render_utils_cuda.so')


_REF_MODS = {}


def load_ref_ext(name='render_utils_cuda'):
    """Import oracle/_ref/<name>.so (the reference's own pybind module)."""
    if name not in _REF_MODS:
        path = os.path.join(HERE, '_ref', name + '.so')
        if not os.path.exists(path):
            raise FileNotFoundError(f'{path} missing: run `python oracle/build_ref.py` where /root/reference exists')
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _REF_MODS[name] = mod
    return _REF_MODS[name]


class RefExtOps:
    """The reference's own CUDA kernels (GPU box only).  Same interface as CpuOps.

    The reference allocates outputs with ``.device(torch::kCUDA)`` and launches on the legacy
    default stream (render_utils_kernel.cu:108,213,226), so callers must be on the default stream.
    """
    device = 'cuda'

    def __init__(self):
        self.m = load_ref_ext()

    def cumdist_thres(self, dist, thres):
        """ub360_utils_cuda.cumdist_thres -- the reference's own kernel (lib/dcvgo.py:284)."""
        return load_ref_ext('ub360_utils_cuda').cumdist_thres(dist, thres)

    def __getattr__(self, name):
        return getattr(self.m, name)

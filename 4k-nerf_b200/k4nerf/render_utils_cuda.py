"""Drop-in for the reference's pybind module ``render_utils_cuda`` (lib/cuda/render_utils.cpp:170-184):
the same 13 function names, argument order and return lists, backed by csrc/k4_ops.cu through the
C ABI.  Results are bit-identical to the reference extension (tests/test_gpu_ops_module.py).

Contract mirrored from the reference: inputs must be CUDA and contiguous (``CHECK_INPUT``,
render_utils.cpp:46-48 -> RuntimeError), outputs are freshly allocated tensors.  Unlike the
reference, kernels run on the CURRENT stream and the only host sync is the unavoidable one of
``sample_pts_on_rays`` (the number of points sizes the outputs: the reference's ``.item()``,
render_utils_kernel.cu:212).
"""
import ctypes as C

import torch

from . import _lib

_L = _lib.lib


def _chk(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError('x must be a CUDA tensor')
        if not t.is_contiguous():
            raise RuntimeError('x must be contiguous')


def _s(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr())


def _call(name, *args):
    _lib.check(getattr(_L, name)(*args), name)


def infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far):
    _chk(rays_o, rays_d, xyz_min, xyz_max)
    n = rays_o.shape[0]
    t_min, t_max = torch.empty(n, device=rays_o.device), torch.empty(n, device=rays_o.device)
    _call('k4_op_infer_t_minmax', _p(rays_o), _p(rays_d), _p(xyz_min), _p(xyz_max), float(near), float(far), n, _p(t_min), _p(t_max), _s(rays_o))
    return [t_min, t_max]


def infer_n_samples(rays_d, t_min, t_max, stepdist):
    _chk(rays_d, t_min, t_max)
    n = t_min.shape[0]
    out = torch.empty(n, dtype=torch.int64, device=rays_d.device)
    _call('k4_op_infer_n_samples', _p(rays_d), _p(t_min), _p(t_max), float(stepdist), n, _p(out), _s(rays_d))
    return out


def infer_ray_start_dir(rays_o, rays_d, t_min):
    _chk(rays_o, rays_d, t_min)
    n = rays_o.shape[0]
    start, rdir = torch.empty_like(rays_o), torch.empty_like(rays_o)
    _call('k4_op_infer_ray_start_dir', _p(rays_o), _p(rays_d), _p(t_min), n, _p(start), _p(rdir), _s(rays_o))
    return [start, rdir]


def sample_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist):
    """render_utils_kernel.cu:196-242.  Returns [rays_pts, mask_outbbox, ray_id, step_id, N_steps, t_min, t_max]."""
    _chk(rays_o, rays_d, xyz_min, xyz_max)
    dev = rays_o.device
    n = rays_o.shape[0]
    t_min, t_max = infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far)
    N_steps = infer_n_samples(rays_d, t_min, t_max, stepdist)
    cum = N_steps.cumsum(0)
    total = int(cum[-1].item()) if n else 0
    ray_id = torch.empty(total, dtype=torch.int64, device=dev)
    step_id = torch.empty(total, dtype=torch.int64, device=dev)
    _call('k4_op_fill_ray_step_ids', _p(cum), n, total, _p(ray_id), _p(step_id), _s(rays_o))
    start, rdir = infer_ray_start_dir(rays_o, rays_d, t_min)
    pts = torch.empty((total, 3), dtype=rays_o.dtype, device=dev)
    mask = torch.empty(total, dtype=torch.bool, device=dev)
    _call('k4_op_sample_pts', _p(start), _p(rdir), _p(xyz_min), _p(xyz_max), _p(ray_id), _p(step_id), float(stepdist), total,
          _p(pts), _p(mask), _s(rays_o))
    return [pts, mask, ray_id, step_id, N_steps, t_min, t_max]


def sample_ndc_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, N_samples):
    _chk(rays_o, rays_d, xyz_min, xyz_max)
    n = rays_o.shape[0]
    pts = torch.empty((n, N_samples, 3), dtype=rays_o.dtype, device=rays_o.device)
    mask = torch.empty((n, N_samples), dtype=torch.bool, device=rays_o.device)
    _call('k4_op_sample_ndc_pts', _p(rays_o), _p(rays_d), _p(xyz_min), _p(xyz_max), int(N_samples), n, _p(pts), _p(mask), _s(rays_o))
    return [pts, mask]


def sample_bg_pts_on_rays(rays_o, rays_d, t_max, bg_preserve, N_samples):
    _chk(rays_o, rays_d, t_max)
    n = rays_o.shape[0]
    pts = torch.empty((n, N_samples, 3), dtype=rays_o.dtype, device=rays_o.device)
    _call('k4_op_sample_bg_pts', _p(rays_o), _p(rays_d), _p(t_max), float(bg_preserve), int(N_samples), n, _p(pts), _s(rays_o))
    return pts


def maskcache_lookup(world, xyz, xyz2ijk_scale, xyz2ijk_shift):
    _chk(world, xyz, xyz2ijk_scale, xyz2ijk_shift)
    assert world.dim() == 3 and xyz.dim() == 2 and xyz.shape[1] == 3
    n = xyz.shape[0]
    out = torch.zeros(n, dtype=torch.bool, device=xyz.device)
    if n:
        _call('k4_op_maskcache_lookup', _p(world), _p(xyz), _p(out), _p(xyz2ijk_scale), _p(xyz2ijk_shift),
              world.shape[0], world.shape[1], world.shape[2], n, _s(xyz))
    return out


def raw2alpha(density, shift, interval):
    _chk(density)
    assert density.dim() == 1
    exp_d, alpha = torch.empty_like(density), torch.empty_like(density)
    _call('k4_op_raw2alpha', _p(density), float(shift), float(interval), None, density.shape[0], _p(exp_d), _p(alpha), _s(density))
    return [exp_d, alpha]


def raw2alpha_nonuni(density, shift, interval):
    _chk(density, interval)
    assert density.dim() == 1
    exp_d, alpha = torch.empty_like(density), torch.empty_like(density)
    _call('k4_op_raw2alpha', _p(density), float(shift), 0.0, _p(interval), density.shape[0], _p(exp_d), _p(alpha), _s(density))
    return [exp_d, alpha]


def raw2alpha_backward(exp, grad_back, interval):
    _chk(exp, grad_back)
    grad = torch.empty_like(exp)
    _call('k4_op_raw2alpha_backward', _p(exp), _p(grad_back), float(interval), None, exp.shape[0], _p(grad), _s(exp))
    return grad


def raw2alpha_nonuni_backward(exp, grad_back, interval):
    _chk(exp, grad_back, interval)
    grad = torch.empty_like(exp)
    _call('k4_op_raw2alpha_backward', _p(exp), _p(grad_back), 0.0, _p(interval), exp.shape[0], _p(grad), _s(exp))
    return grad


def alpha2weight(alpha, ray_id, n_rays):
    _chk(alpha, ray_id)
    assert alpha.dim() == 1 and ray_id.dim() == 1 and alpha.shape[0] == ray_id.shape[0]
    dev = alpha.device
    weight, T = torch.zeros_like(alpha), torch.ones_like(alpha)
    last = torch.ones(n_rays, dtype=alpha.dtype, device=dev)
    i_start = torch.zeros(n_rays, dtype=torch.int64, device=dev)
    i_end = torch.zeros(n_rays, dtype=torch.int64, device=dev)
    _call('k4_op_alpha2weight', _p(alpha), _p(ray_id), int(n_rays), alpha.shape[0], _p(weight), _p(T), _p(last), _p(i_start), _p(i_end), _s(alpha))
    return [weight, T, last, i_start, i_end]


def alpha2weight_backward(alpha, weight, T, alphainv_last, i_start, i_end, n_rays, grad_weights, grad_last):
    _chk(alpha, weight, T, alphainv_last, i_start, i_end, grad_weights, grad_last)
    grad = torch.zeros_like(alpha)
    if n_rays:
        _call('k4_op_alpha2weight_backward', _p(alpha), _p(weight), _p(T), _p(alphainv_last), _p(i_start), _p(i_end), int(n_rays),
              _p(grad_weights), _p(grad_last), _p(grad), _s(alpha))
    return grad


def cumdist_thres(dist, thres):
    """``ub360_utils_cuda.cumdist_thres`` (lib/cuda/ub360_utils.cpp:20-22), the one function of the
    reference's second sampling extension: dist [n_rays, n_pts] -> bool mask of the kept samples."""
    _chk(dist)
    mask = torch.zeros(dist.shape, dtype=torch.bool, device=dist.device)
    if dist.numel():
        _call('k4_op_cumdist_thres', _p(dist), float(thres), dist.shape[0], dist.shape[1], _p(mask), _s(dist))
    return mask

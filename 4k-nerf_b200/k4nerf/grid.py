"""Parameter holders with the reference's module tree and state-dict keys (lib/grid.py).

At inference the reference's DenseGrid.forward / MaskGrid.forward (lib/grid.py:117-128,295-304)
never run as separate ops: interpolation and occupancy lookup happen inside the fused marcher
(csrc/k4_march.cu).  The ``forward`` methods here exist for the training-side callers (grid
maintenance, the op-level training forward); the classes own the tensors so that reference
checkpoints load unchanged: ``density.grid``, ``density.xyz_min``, ``density.xyz_max``, ``mask_cache.mask``,
``mask_cache.xyz2ijk_scale``, ``mask_cache.xyz2ijk_shift``.
"""
import torch
import torch.nn as nn


def create_grid(type, **kwargs):
    if type == 'DenseGrid':
        return DenseGrid(**kwargs)
    # TensoRFGrid / VQGrid: no shipped config selects them (configs/default.py:85-86) -- out of scope
    raise NotImplementedError(f'k4nerf supports density_type/k0_type "DenseGrid" only, got {type!r}')


class _GridSample(torch.autograd.Function):
    """DenseGrid.forward on the library's fused kernels (csrc/k4_train.cu): forward = normalise + trilinear gather +
    transpose in one launch, backward = gradient scatter into the grid (the sample positions carry no gradient in the
    reference either: ray_pts come out of the no-grad samplers)."""

    @staticmethod
    def forward(ctx, grid, xyz, lo, hi):
        import ctypes as C
        from . import _lib
        from .render_utils_cuda import _p, _s, _call
        xyz = xyz.detach().reshape(-1, 3).to(torch.float32).contiguous()
        g = grid.detach().contiguous()
        Cn, X, Y, Z = (int(v) for v in g.shape[1:])
        out = torch.empty((xyz.shape[0], Cn), device=g.device, dtype=torch.float32)
        lo_c, hi_c = (C.c_float * 3)(*lo), (C.c_float * 3)(*hi)
        with torch.cuda.device(g.device):
            _call('k4_op_grid_sample', _p(g), Cn, X, Y, Z, lo_c, hi_c, _p(xyz), xyz.shape[0], _p(out), _s(g))
        ctx.save_for_backward(xyz)
        ctx.meta = (tuple(grid.shape), lo, hi)
        del _lib
        return out

    @staticmethod
    def backward(ctx, grad_out):
        import ctypes as C
        from .render_utils_cuda import _p, _s, _call
        (xyz,) = ctx.saved_tensors
        shape, lo, hi = ctx.meta
        Cn, X, Y, Z = (int(v) for v in shape[1:])
        grad_grid = torch.zeros(shape, device=grad_out.device, dtype=torch.float32)
        go = grad_out.to(torch.float32).contiguous()
        lo_c, hi_c = (C.c_float * 3)(*lo), (C.c_float * 3)(*hi)
        with torch.cuda.device(go.device):
            _call('k4_op_grid_sample_backward', _p(go), Cn, X, Y, Z, lo_c, hi_c, _p(xyz), xyz.shape[0], _p(grad_grid), _s(go))
        return grad_grid, None, None, None


class DenseGrid(nn.Module):
    def __init__(self, channels, world_size, xyz_min, xyz_max, **kwargs):
        super().__init__()
        self.channels = channels
        self.world_size = world_size
        self.register_buffer('xyz_min', torch.as_tensor(xyz_min, dtype=torch.float32).clone())
        self.register_buffer('xyz_max', torch.as_tensor(xyz_max, dtype=torch.float32).clone())
        self.grid = nn.Parameter(torch.zeros([1, channels, *[int(w) for w in world_size]]))

    def _host_box(self):
        """xyz_min / xyz_max as host floats, cached by the buffers' versions (no per-call device sync)."""
        key = (self.xyz_min.data_ptr(), self.xyz_min._version, self.xyz_max.data_ptr(), self.xyz_max._version)
        c = self.__dict__.get('_k4_box')
        if c is None or c[0] != key:
            c = (key, tuple(self.xyz_min.detach().cpu().tolist()), tuple(self.xyz_max.detach().cpu().tolist()))
            self.__dict__['_k4_box'] = c
        return c[1], c[2]

    def forward(self, xyz):
        """Trilinear lookup with autograd (lib/grid.py:117-128).  CUDA fp32 grids whose sample positions carry no
        gradient (every caller in the reference) run on the library's fused forward / gradient-scatter kernels
        (k4_op_grid_sample[_backward]); anything else takes the reference's ATen grid_sample path."""
        import torch.nn.functional as F
        shape = xyz.shape[:-1]
        if (self.grid.is_cuda and self.grid.dtype == torch.float32 and xyz.is_cuda and not xyz.requires_grad
                and self.channels > 0 and self.grid.numel() < (1 << 31) * self.channels):
            lo, hi = self._host_box()
            out = _GridSample.apply(self.grid, xyz, lo, hi).reshape(*shape, self.channels)
            return out.squeeze(-1) if self.channels == 1 else out
        xyz = xyz.reshape(1, 1, 1, -1, 3)
        ind_norm = ((xyz - self.xyz_min) / (self.xyz_max - self.xyz_min)).flip((-1,)) * 2 - 1
        out = F.grid_sample(self.grid, ind_norm, mode='bilinear', align_corners=True)
        out = out.reshape(self.channels, -1).T.reshape(*shape, self.channels)
        return out.squeeze(-1) if self.channels == 1 else out

    @torch.no_grad()
    def scale_volume_grid(self, new_world_size):
        """lib/grid.py:130-135: trilinear resample (align_corners=True) to the new resolution, one launch
        of csrc/k4_train.cu for all channels."""
        from . import _lib
        from .render_utils_cuda import _p, _s, _call
        ws = [int(w) for w in new_world_size]
        self.world_size = new_world_size
        if self.channels == 0:
            self.grid = nn.Parameter(torch.zeros([1, 0, *ws], device=self.grid.device))
            return
        src = self.grid.data.contiguous()
        dst = torch.empty([1, self.channels, *ws], device=src.device, dtype=torch.float32)
        _call('k4_op_resample_trilinear', _p(src), self.channels, int(src.shape[2]), int(src.shape[3]), int(src.shape[4]),
              _p(dst), ws[0], ws[1], ws[2], _s(src))
        self.grid = nn.Parameter(dst)
        del _lib

    def total_variation_add_grad(self, wx, wy, wz, dense_mode):
        """Add the total-variation gradient to ``grid.grad`` in place (lib/grid.py:137-140)."""
        from . import total_variation_cuda
        total_variation_cuda.total_variation_add_grad(self.grid, self.grid.grad, float(wx), float(wy), float(wz), dense_mode)

    def get_dense_grid(self):
        return self.grid

    def extra_repr(self):
        return f'channels={self.channels}, world_size={[int(w) for w in self.world_size]}'


class MaskGrid(nn.Module):
    """Occupancy grid; xyz2ijk arithmetic of lib/grid.py:291-293."""

    def __init__(self, path=None, mask_cache_thres=None, mask=None, xyz_min=None, xyz_max=None):
        super().__init__()
        if path is not None:
            # occupancy of a coarse-stage checkpoint (lib/grid.py:277-285): dilate the density by one voxel,
            # alpha with the softplus form of the activation, threshold
            import torch.nn.functional as F
            st = torch.load(path, map_location='cpu', weights_only=False)
            sd, kw = st['model_state_dict'], st['model_kwargs']
            self.mask_cache_thres = mask_cache_thres
            dens = F.max_pool3d(sd['density.grid'].float(), kernel_size=3, padding=1, stride=1)
            alpha = 1 - torch.exp(-F.softplus(dens + sd['act_shift'].float()) * kw['voxel_size_ratio'])
            mask = (alpha >= mask_cache_thres)[0, 0]
            xyz_min, xyz_max = kw['xyz_min'], kw['xyz_max']
        mask = mask.bool()
        xyz_min = torch.as_tensor(xyz_min, dtype=torch.float32).cpu()
        xyz_max = torch.as_tensor(xyz_max, dtype=torch.float32).cpu()
        self.register_buffer('mask', mask)
        xyz_len = xyz_max - xyz_min
        self.register_buffer('xyz2ijk_scale', (torch.Tensor(list(mask.shape)) - 1) / xyz_len)
        self.register_buffer('xyz2ijk_shift', -xyz_min * self.xyz2ijk_scale)

    @torch.no_grad()
    def forward(self, xyz):
        """Nearest-voxel occupancy of world points (lib/grid.py:295-304)."""
        from . import render_utils_cuda
        shape = xyz.shape[:-1]
        xyz = xyz.reshape(-1, 3).contiguous()
        return render_utils_cuda.maskcache_lookup(self.mask, xyz, self.xyz2ijk_scale, self.xyz2ijk_shift).reshape(shape)

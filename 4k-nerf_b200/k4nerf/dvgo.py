"""DirectVoxGO with the reference constructor / checkpoint / forward contract (lib/dvgo.py:23-448),
rendered by the fused sm_100a marcher instead of the reference's op-by-op pipeline."""
import numpy as np
import torch
import torch.nn as nn

from . import _lib, grid
from ._scene import FusedRenderMixin, cached_host, host_float
from .maintain import GridMaintenanceMixin
from .coarse import CoarseStageMixin


class DirectVoxGO(FusedRenderMixin, GridMaintenanceMixin, CoarseStageMixin, nn.Module):
    _k4_kind = _lib.K4_KIND_DVGO

    def __init__(self, xyz_min, xyz_max,
                 num_voxels=0, num_voxels_base=0,
                 alpha_init=None,
                 mask_cache_path=None, mask_cache_thres=1e-3, mask_cache_world_size=None,
                 fast_color_thres=0,
                 density_type='DenseGrid', k0_type='DenseGrid',
                 density_config={}, k0_config={},
                 rgbnet_dim=0, rgbnet_direct=False, rgbnet_full_implicit=False,
                 rgbnet_depth=3, rgbnet_width=128,
                 viewbase_pe=4,
                 **kwargs):
        super().__init__()
        if rgbnet_full_implicit:
            raise NotImplementedError('rgbnet_full_implicit is not used by any shipped config')
        self.register_buffer('xyz_min', torch.Tensor(xyz_min))
        self.register_buffer('xyz_max', torch.Tensor(xyz_max))
        self.fast_color_thres = fast_color_thres
        self.num_voxels_base = num_voxels_base
        self.voxel_size_base = ((self.xyz_max - self.xyz_min).prod() / self.num_voxels_base).pow(1 / 3)
        self.alpha_init = alpha_init
        self.register_buffer('act_shift', torch.FloatTensor([np.log(1 / (1 - alpha_init) - 1)]))
        self._set_grid_resolution(num_voxels)

        self.density_type, self.density_config = density_type, density_config
        self.k0_type, self.k0_config = k0_type, k0_config
        self.density = grid.create_grid(density_type, channels=1, world_size=self.world_size,
                                        xyz_min=self.xyz_min, xyz_max=self.xyz_max, config=density_config)
        self.rgbnet_kwargs = {
            'rgbnet_dim': rgbnet_dim, 'rgbnet_direct': rgbnet_direct,
            'rgbnet_full_implicit': rgbnet_full_implicit,
            'rgbnet_depth': rgbnet_depth, 'rgbnet_width': rgbnet_width,
            'viewbase_pe': viewbase_pe,
        }
        self.rgbnet_full_implicit = rgbnet_full_implicit
        self.rgbnet_direct = rgbnet_direct
        self.viewbase_pe = viewbase_pe
        self.dim_rend, self.act_type, self.mode_type = 3, 'mlp', 'mlp'
        if rgbnet_dim <= 0:
            self.k0_dim = 3                       # coarse stage: colour grid, no MLP
            self.rgbnet = None
        else:
            self.k0_dim = rgbnet_dim
            self.register_buffer('viewfreq', torch.FloatTensor([(2 ** i) for i in range(viewbase_pe)]))
            dim0 = 3 + 3 * viewbase_pe * 2 + (self.k0_dim if rgbnet_direct else self.k0_dim - 3)
            self.dim0 = dim0
            self.rgbnet = nn.Sequential(
                nn.Linear(dim0, rgbnet_width), nn.ReLU(inplace=True),
                *[nn.Sequential(nn.Linear(rgbnet_width, rgbnet_width), nn.ReLU(inplace=True))
                  for _ in range(rgbnet_depth - 2)],
                nn.Linear(rgbnet_width, 3),
            )
            nn.init.constant_(self.rgbnet[-1].bias, 0)
        self.k0 = grid.create_grid(k0_type, channels=self.k0_dim, world_size=self.world_size,
                                   xyz_min=self.xyz_min, xyz_max=self.xyz_max, config=k0_config)

        self.mask_cache_path = mask_cache_path
        self.mask_cache_thres = mask_cache_thres
        if mask_cache_world_size is None:
            mask_cache_world_size = self.world_size
        if mask_cache_path:
            from .maintain import mask_from_coarse_checkpoint
            mask = mask_from_coarse_checkpoint(mask_cache_path, mask_cache_thres, self.xyz_min, self.xyz_max, mask_cache_world_size)
        else:
            mask = torch.ones([int(w) for w in mask_cache_world_size], dtype=torch.bool)
        self.mask_cache = grid.MaskGrid(path=None, mask=mask, xyz_min=self.xyz_min, xyz_max=self.xyz_max)

    def _set_grid_resolution(self, num_voxels):
        # lib/dvgo.py:152-158
        self.num_voxels = num_voxels
        self.voxel_size = ((self.xyz_max - self.xyz_min).prod() / num_voxels).pow(1 / 3)
        self.world_size = ((self.xyz_max - self.xyz_min) / self.voxel_size).long()
        self.max_world_size = self.world_size.max()
        self.voxel_size_ratio = self.voxel_size / self.voxel_size_base

    def get_kwargs(self):
        return {
            'xyz_min': self.xyz_min.cpu().numpy(), 'xyz_max': self.xyz_max.cpu().numpy(),
            'num_voxels': self.num_voxels, 'num_voxels_base': self.num_voxels_base,
            'alpha_init': self.alpha_init, 'voxel_size_ratio': self.voxel_size_ratio,
            'mask_cache_path': self.mask_cache_path, 'mask_cache_thres': self.mask_cache_thres,
            'mask_cache_world_size': list(self.mask_cache.mask.shape),
            'fast_color_thres': self.fast_color_thres,
            'density_type': self.density_type, 'k0_type': self.k0_type,
            'density_config': self.density_config, 'k0_config': self.k0_config,
            'mode_type': self.mode_type, 'act_type': self.act_type, 'dim_rend': self.dim_rend,
            **self.rgbnet_kwargs,
        }

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # checkpoints may hold grids of a different (progressively scaled) resolution than the
        # constructor arithmetic yields ("world-size arithmetic"): adopt the
        # checkpoint's shapes instead of failing.
        for name, mod in (('density', self.density), ('k0', self.k0)):
            key = prefix + name + '.grid'
            if key in state_dict and state_dict[key].shape != mod.grid.shape:
                mod.grid = nn.Parameter(torch.zeros_like(state_dict[key]))
        key = prefix + 'mask_cache.mask'
        if key in state_dict and state_dict[key].shape != self.mask_cache.mask.shape:
            self.mask_cache.mask = torch.zeros_like(state_dict[key])
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        # the constructor arithmetic may be off by one voxel from the checkpoint (fp32 cube root):
        # keep world_size / max_world_size (depth normalisation, TV weights, scale_volume_grid) consistent with the grids
        self.world_size = torch.tensor(list(self.density.grid.shape[2:]), dtype=torch.long)
        self.max_world_size = self.world_size.max()
        self.invalidate_scene()

    def _scene_extra(self):
        return {
            'act_shift': cached_host(self, 'act_shift', [self.act_shift], lambda: float(self.act_shift)),
            'voxel_size': host_float(self, 'voxel_size'),
            'rgbnet_direct': self.rgbnet_direct, 'viewbase_pe': self.viewbase_pe if self.rgbnet is not None else 0,
        }


# --- ray helpers with the reference names (lib/dvgo.py:516-582); generated on the device -------------
def get_rays_of_a_view(H, W, K, c2w, ndc, inverse_y, flip_x, flip_y, mode='center', rows=None, device=None):
    """Pixel-centre rays of one view, ``[H, W, 3]`` each, produced by ``k4_make_rays``.
    ``rows`` (int32 CUDA tensor of image-row indices, not in the reference signature): only those rows are
    generated, ``[len(rows), W, 3]`` -- a rank's share of the frame in the multi-GPU driver."""
    if mode != 'center':
        # 'lefttop' / 'random' sub-pixel positions are training-data options (lib/dvgo.py:520-528): plain torch
        from .coarse import get_rays_of_a_view_torch
        if rows is not None:
            raise NotImplementedError("rows= (multi-GPU frame sharding) is built for mode='center' only")
        c2w_t = torch.as_tensor(c2w, dtype=torch.float32)
        if device is not None:
            c2w_t = c2w_t.to(device)
        return get_rays_of_a_view_torch(H, W, K, c2w_t, ndc, inverse_y, flip_x, flip_y, mode=mode)
    import ctypes as C
    c2w = torch.as_tensor(c2w, dtype=torch.float32)
    if device is not None:
        dev = torch.device(device)
    elif rows is not None:
        dev = rows.device
    else:
        dev = c2w.device if c2w.is_cuda else torch.device('cuda', torch.cuda.current_device())
    Kh = (C.c_float * 9)(*np.asarray(K, dtype=np.float32).reshape(-1)[:9].tolist())
    ch = (C.c_float * 12)(*c2w.detach().cpu().reshape(-1)[:12].tolist())
    H, W = int(H), int(W)
    n_rows = H if rows is None else int(rows.numel())
    if rows is not None:
        assert rows.is_cuda and rows.dtype == torch.int32 and rows.is_contiguous()
    out = [torch.empty((n_rows, W, 3), device=dev, dtype=torch.float32) for _ in range(3)]
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib.k4_make_rays_rows(Kh, ch, H, W, int(bool(ndc)), int(bool(inverse_y)), int(bool(flip_x)),
                                              int(bool(flip_y)), rows.data_ptr() if rows is not None else None, n_rows,
                                              out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                              C.c_void_p(stream)), 'k4_make_rays_rows')
    return out[0], out[1], out[2]


# training-ray helpers with the reference's names (lib/dvgo.py:516-544,585-697,761-768), plain torch on top of the above
from .coarse import (get_rays, get_training_rays, get_training_rays_flatten,          # noqa: E402,F401
                     get_training_rays_in_maskcache_sampling, batch_indices_generator)

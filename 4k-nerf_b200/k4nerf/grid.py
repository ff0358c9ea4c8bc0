"""Parameter holders with the reference's module tree and state-dict keys (lib/grid.py).

The reference's DenseGrid.forward / MaskGrid.forward (lib/grid.py:117-128,295-304) are not
reimplemented as separate ops: interpolation and occupancy lookup happen inside the fused marcher
(csrc/k4_march.cu).  These classes only own the tensors so that reference checkpoints load
unchanged: ``density.grid``, ``density.xyz_min``, ``density.xyz_max``, ``mask_cache.mask``,
``mask_cache.xyz2ijk_scale``, ``mask_cache.xyz2ijk_shift``.
"""
import torch
import torch.nn as nn


def create_grid(type, **kwargs):
    if type == 'DenseGrid':
        return DenseGrid(**kwargs)
    # TensoRFGrid / VQGrid: no shipped config selects them (configs/default.py:85-86) -- out of scope
    raise NotImplementedError(f'k4nerf supports density_type/k0_type "DenseGrid" only, got {type!r}')


class DenseGrid(nn.Module):
    def __init__(self, channels, world_size, xyz_min, xyz_max, **kwargs):
        super().__init__()
        self.channels = channels
        self.world_size = world_size
        self.register_buffer('xyz_min', torch.as_tensor(xyz_min, dtype=torch.float32).clone())
        self.register_buffer('xyz_max', torch.as_tensor(xyz_max, dtype=torch.float32).clone())
        self.grid = nn.Parameter(torch.zeros([1, channels, *[int(w) for w in world_size]]))

    def get_dense_grid(self):
        return self.grid

    def extra_repr(self):
        return f'channels={self.channels}, world_size={[int(w) for w in self.world_size]}'


class MaskGrid(nn.Module):
    """Occupancy grid; xyz2ijk arithmetic of lib/grid.py:291-293."""

    def __init__(self, path=None, mask_cache_thres=None, mask=None, xyz_min=None, xyz_max=None):
        super().__init__()
        if path is not None:
            raise NotImplementedError('mask_cache_path (coarse-stage checkpoint) is a training-time feature')
        mask = mask.bool()
        xyz_min = torch.as_tensor(xyz_min, dtype=torch.float32).cpu()
        xyz_max = torch.as_tensor(xyz_max, dtype=torch.float32).cpu()
        self.register_buffer('mask', mask)
        xyz_len = xyz_max - xyz_min
        self.register_buffer('xyz2ijk_scale', (torch.Tensor(list(mask.shape)) - 1) / xyz_len)
        self.register_buffer('xyz2ijk_shift', -xyz_min * self.xyz2ijk_scale)

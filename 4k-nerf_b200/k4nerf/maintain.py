"""Grid-maintenance steps of the reference models, shared by DirectVoxGO,
DirectMPIGO and DirectContractedVoxGO: ``activate_density``, ``update_occupancy_cache``,
``scale_volume_grid`` and the two total-variation hooks.  The kernels are in csrc/k4_train.cu; every
method ends by invalidating the packed scene so the next render repacks the changed tensors."""
import ctypes as C

import numpy as np
import torch

from . import _lib, grid, render_utils_cuda
from .render_utils_cuda import _p, _s, _call


def _f3(t):
    return (C.c_float * 3)(*[float(x) for x in t.detach().cpu().tolist()])


def _linspaces(xyz_min, xyz_max, shape, device):
    # the reference builds these with torch.linspace on its default (CUDA) device
    return [torch.linspace(float(xyz_min[a]), float(xyz_max[a]), int(shape[a]), device=device) for a in range(3)]


def mask_from_coarse_checkpoint(path, thres, xyz_min, xyz_max, mask_world_size):
    """Initial occupancy of a fine-stage model: the coarse checkpoint's occupancy (MaskGrid(path=...))
    sampled at this model's mask-grid points (lib/dvgo.py:134-145, lib/dmpigo.py:140-151).  The lookup is
    the CUDA op, as in the reference (which builds these on its default CUDA device)."""
    if not torch.cuda.is_available():
        raise RuntimeError('mask_cache_path needs a CUDA device (maskcache_lookup has no CPU path)')
    dev = torch.device('cuda', torch.cuda.current_device())
    coarse = grid.MaskGrid(path=path, mask_cache_thres=thres).to(dev)
    lx, ly, lz = _linspaces(xyz_min, xyz_max, mask_world_size, dev)
    return coarse(torch.stack(torch.meshgrid(lx, ly, lz, indexing='ij'), -1)).cpu()


class GridMaintenanceMixin:
    _k4_shift_in_alpha = True        # DirectMPIGO: activate_density uses shift 0 (lib/dmpigo.py:258-261)

    def _alpha_shift(self):
        return float(self.act_shift) if self._k4_shift_in_alpha else 0.0

    @torch.no_grad()
    def activate_density(self, density, interval=None):
        """lib/dvgo.py:276-279 / lib/dmpigo.py:258-261 (values only; the autograd version is
        render_utils_cuda.raw2alpha + raw2alpha_backward)."""
        interval = interval if interval is not None else self.voxel_size_ratio
        shape = density.shape
        return render_utils_cuda.raw2alpha(density.detach().flatten().contiguous(), self._alpha_shift(), float(interval))[1].reshape(shape)

    @torch.no_grad()
    def update_occupancy_cache(self):
        """mask_cache.mask &= max_pool3d(alpha(density at the mask grid's points)) > fast_color_thres
        (lib/dvgo.py:224-233, lib/dmpigo.py:212-224, lib/dcvgo.py:177-190): two launches, no
        intermediate point list."""
        mask = self.mask_cache.mask
        dev = mask.device
        g = self.density.grid
        lx, ly, lz = _linspaces(self.xyz_min, self.xyz_max, mask.shape, dev)
        alpha = torch.empty(mask.shape, device=dev, dtype=torch.float32)
        mn, mx = _f3(self.density.xyz_min), _f3(self.density.xyz_max)
        _call('k4_op_grid_alpha', _p(g), int(g.shape[2]), int(g.shape[3]), int(g.shape[4]), mn, mx, _p(lx), _p(ly), _p(lz),
              int(mask.shape[0]), int(mask.shape[1]), int(mask.shape[2]), self._alpha_shift(), float(self.voxel_size_ratio),
              _p(alpha), _s(mask))
        _call('k4_op_maxpool3_thres_and', _p(alpha), int(mask.shape[0]), int(mask.shape[1]), int(mask.shape[2]),
              float(self.fast_color_thres), _p(mask), _s(mask))
        self.invalidate_scene()

    def _rescaled_density_for_mask(self):
        return self.density.get_dense_grid()

    @torch.no_grad()
    def scale_volume_grid(self, num_voxels, *res_args):
        """Resample density / k0 to the new resolution and rebuild the occupancy grid at that resolution
        (lib/dvgo.py:200-221, lib/dmpigo.py:189-211, lib/dcvgo.py:155-174)."""
        self._set_grid_resolution(num_voxels, *res_args)
        self.density.scale_volume_grid(self.world_size)
        self.k0.scale_volume_grid(self.world_size)
        ws = [int(w) for w in self.world_size.tolist()]
        if np.prod(ws) <= 256 ** 3:
            dev = self.density.grid.device
            lx, ly, lz = _linspaces(self.xyz_min, self.xyz_max, ws, dev)
            pts = torch.stack(torch.meshgrid(lx, ly, lz, indexing='ij'), -1)
            mask = self.mask_cache(pts).contiguous()
            alpha = self.activate_density(self._rescaled_density_for_mask())[0, 0].contiguous()
            _call('k4_op_maxpool3_thres_and', _p(alpha), ws[0], ws[1], ws[2], float(self.fast_color_thres), _p(mask), _s(mask))
            self.mask_cache = grid.MaskGrid(path=None, mask=mask, xyz_min=self.xyz_min, xyz_max=self.xyz_max).to(dev)
        self.invalidate_scene()

    def _tv_weights(self, weight):
        w = weight * self.world_size.max() / 128             # lib/dvgo.py:268-274
        return w, w, w

    def density_total_variation_add_grad(self, weight, dense_mode):
        self.density.total_variation_add_grad(*self._tv_weights(weight), dense_mode)

    def k0_total_variation_add_grad(self, weight, dense_mode):
        self.k0.total_variation_add_grad(*self._tv_weights(weight), dense_mode)

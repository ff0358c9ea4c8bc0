"""Coarse-stage preparation and training-ray helpers of the reference (steps either side of the render path during
training), on top of the op-level kernels:

  DirectVoxGO.maskout_near_cam_vox     lib/dvgo.py:185-198
  DirectVoxGO.voxel_count_views        lib/dvgo.py:235-266   (per-voxel view counts -> per-voxel learning rates, run_sr.py:442)
  DirectVoxGO.hit_coarse_geo           lib/dvgo.py:281-293
  DirectVoxGO.sample_ray               lib/dvgo.py:295-325
  DirectMPIGO / DirectContractedVoxGO.update_occupancy_cache_lt_nviews   lib/dmpigo.py:228-246, lib/dcvgo.py:192-210
  get_rays (all three pixel modes), get_training_rays, get_training_rays_flatten,
  get_training_rays_in_maskcache_sampling, batch_indices_generator       lib/dvgo.py:516-544,585-697

The view counts are the reference's trick "back-propagate sum(trilinear weights) into a grid of ones and look at the
gradient": here the gradient-scatter kernel (k4_op_grid_sample_backward) is called directly with a gradient of ones,
no autograd graph and no throw-away DenseGrid module per view.
"""
import ctypes as C

import numpy as np
import torch

from . import render_utils_cuda as ops
from .render_utils_cuda import _p, _s, _call


def _touch_weights(acc, pts, lo, hi):
    """acc [1,1,X,Y,Z] += sum over points of their trilinear corner weights (== d/dgrid of sum(grid(pts)))."""
    pts = pts.reshape(-1, 3).to(torch.float32).contiguous()
    if pts.shape[0] == 0:
        return
    ones = torch.ones((pts.shape[0], 1), device=pts.device, dtype=torch.float32)
    X, Y, Z = (int(v) for v in acc.shape[2:])
    with torch.cuda.device(acc.device):
        _call('k4_op_grid_sample_backward', _p(ones), 1, X, Y, Z, (C.c_float * 3)(*lo), (C.c_float * 3)(*hi), _p(pts), pts.shape[0],
              _p(acc), _s(acc))


class CoarseStageMixin:
    """Methods of the reference models that prepare the grids before / between training stages."""

    def _box(self):
        return self.density._host_box()

    @torch.no_grad()
    def maskout_near_cam_vox(self, cam_o, near_clip):
        """Density -100 at grid points closer than `near_clip` to any camera centre (lib/dvgo.py:185-198)."""
        dev = self.density.grid.device
        lo, hi = self._box()
        ax = [torch.linspace(lo[a], hi[a], int(self.world_size[a]), device=dev) for a in range(3)]
        pts = torch.stack(torch.meshgrid(*ax, indexing='ij'), -1)
        nearest = torch.full(pts.shape[:-1], float('inf'), device=dev)
        for block in torch.as_tensor(cam_o, dtype=torch.float32, device=dev).split(100):      # bounded memory, as the reference
            nearest = torch.minimum(nearest, (pts.unsqueeze(-2) - block).pow(2).sum(-1).sqrt().amin(-1))
        self.density.grid[nearest[None, None] <= near_clip] = -100
        self.invalidate_scene()

    def sample_ray(self, rays_o, rays_d, near, far, stepsize, **render_kwargs):
        """lib/dvgo.py:295-325: in-box sample points of DirectVoxGO rays with their ray / step ids."""
        from . import _lib
        if self._k4_kind != _lib.K4_KIND_DVGO:
            raise NotImplementedError('sample_ray of DirectMPIGO / DirectContractedVoxGO: see k4nerf.train_forward._SAMPLERS '
                                      '(the fused kernel samples in-kernel; the materialised samplers live with the training forward)')
        far = 1e9
        stepdist = stepsize * self.voxel_size
        pts, outside, ray_id, step_id, *_ = ops.sample_pts_on_rays(
            rays_o.contiguous(), rays_d.contiguous(), self.xyz_min, self.xyz_max, near, far, stepdist)
        keep = ~outside
        return pts[keep], ray_id[keep], step_id[keep]

    @torch.no_grad()
    def hit_coarse_geo(self, rays_o, rays_d, near, far, stepsize, **render_kwargs):
        """Which rays pass through at least one occupied voxel of the (coarse) occupancy grid (lib/dvgo.py:281-293)."""
        shape = rays_o.shape[:-1]
        pts, ray_id, _ = self.sample_ray(rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), near, far, stepsize)
        hit = torch.zeros(int(np.prod(shape)), dtype=torch.bool, device=pts.device)
        hit[ray_id[self.mask_cache(pts)]] = True
        return hit.reshape(shape)

    def voxel_count_views(self, rays_o_tr, rays_d_tr, imsz, near, far, stepsize, downrate=1, irregular_shape=False):
        """In how many training views is a voxel touched by samples of total trilinear weight > 1 (lib/dvgo.py:235-266).
        Sampling as the reference: ray/box slab test, t_min + k * stepsize * voxel_size / |d| for k < N_samples (points
        beyond the box fall outside the grid and add nothing)."""
        dev = self.density.grid.device
        lo, hi = self._box()
        n_samples = int(np.linalg.norm(np.array(self.world_size.cpu()) + 1) / stepsize) + 1
        k = torch.arange(n_samples, device=dev)[None].float()
        count = torch.zeros_like(self.density.get_dense_grid())
        far = 1e9
        for view_o, view_d in zip(rays_o_tr.split(imsz), rays_d_tr.split(imsz)):
            if not irregular_shape:
                view_o = view_o[::downrate, ::downrate].flatten(0, -2)
                view_d = view_d[::downrate, ::downrate].flatten(0, -2)
            acc = torch.zeros_like(count)
            for o, d in zip(view_o.to(dev).split(10000), view_d.to(dev).split(10000)):
                v = torch.where(d == 0, torch.full_like(d, 1e-6), d)
                a, b = (self.xyz_max - o) / v, (self.xyz_min - o) / v
                t0 = torch.minimum(a, b).amax(-1).clamp(min=near, max=far)
                t = t0[..., None] + stepsize * self.voxel_size * k / d.norm(dim=-1, keepdim=True)
                _touch_weights(acc, o[..., None, :] + d[..., None, :] * t[..., None], lo, hi)
            count += (acc > 1)
        return count

    @torch.no_grad()
    def update_occupancy_cache_lt_nviews(self, rays_o_tr, rays_d_tr, imsz, render_kwargs, maskout_lt_nviews):
        """mask &= (#views whose samples touch the voxel) >= maskout_lt_nviews (lib/dmpigo.py:228-246, lib/dcvgo.py:192-210);
        the samples are those of the model's own sampler (train_forward)."""
        from . import train_forward
        dev = self.density.grid.device
        lo, hi = self._box()
        count = torch.zeros_like(self.density.get_dense_grid()).long()
        sampler = train_forward._SAMPLERS[self._k4_kind]
        for view_o, view_d in zip(rays_o_tr.split(imsz), rays_d_tr.split(imsz)):
            acc = torch.zeros(count.shape, device=dev)
            for o, d in zip(view_o.split(8192), view_d.split(8192)):
                _touch_weights(acc, sampler(self, o.to(dev), d.to(dev), render_kwargs)[0], lo, hi)
            count += (acc > 1)
        self.mask_cache.mask &= (count >= maskout_lt_nviews)[0, 0]
        self.invalidate_scene()


# ---------------------------------------------------------------------------------------------------------------
# rays of training views (plain torch: this is data preparation, not the hot path; the render path uses k4_make_rays)
# ---------------------------------------------------------------------------------------------------------------
def get_rays(H, W, K, c2w, inverse_y, flip_x, flip_y, mode='center'):
    """Camera rays through every pixel (lib/dvgo.py:516-544): `mode` picks the sub-pixel position -- 'lefttop' (pixel
    corner), 'center' (+0.5) or 'random' (uniform jitter per pixel)."""
    c2w = torch.as_tensor(c2w, dtype=torch.float32)
    dev = c2w.device
    K = np.asarray(K, dtype=np.float32)
    u, v = torch.meshgrid(torch.linspace(0, W - 1, W, device=dev), torch.linspace(0, H - 1, H, device=dev), indexing='ij')
    u, v = u.t().float(), v.t().float()
    if mode == 'center':
        u, v = u + 0.5, v + 0.5
    elif mode == 'random':
        u, v = u + torch.rand_like(u), v + torch.rand_like(v)
    elif mode != 'lefttop':
        raise NotImplementedError(mode)
    if flip_x:
        u = u.flip((1,))
    if flip_y:
        v = v.flip((0,))
    x, y = (u - K[0][2]) / K[0][0], (v - K[1][2]) / K[1][1]
    dirs = torch.stack([x, y, torch.ones_like(u)], -1) if inverse_y else torch.stack([x, -y, -torch.ones_like(u)], -1)
    rays_d = torch.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1)       # rotate: dot every direction with the rows of R
    rays_o = c2w[:3, 3].expand(rays_d.shape)
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """Rays of a forward-facing scene in normalised device coordinates (lib/dvgo.py:557-575, the NeRF LLFF warp): origins
    moved onto the near plane z = -near, then x, y scaled by the frustum (2 focal / W, 2 focal / H) over -z and z mapped
    to 1 + 2 near / z; directions are the differences to the rays' images at infinity."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    o = rays_o + t[..., None] * rays_d
    sx, sy = -2. * focal / W, -2. * focal / H          # == -1 / (W / (2 focal)), -1 / (H / (2 focal))
    ox, oy, oz = o[..., 0] / o[..., 2], o[..., 1] / o[..., 2], o[..., 2]
    new_o = torch.stack([sx * ox, sy * oy, 1. + 2. * near / oz], -1)
    new_d = torch.stack([sx * (rays_d[..., 0] / rays_d[..., 2] - ox), sy * (rays_d[..., 1] / rays_d[..., 2] - oy),
                         -2. * near / oz], -1)
    return new_o, new_d


def get_rays_of_a_view_torch(H, W, K, c2w, ndc, inverse_y, flip_x, flip_y, mode='center'):
    """lib/dvgo.py:577-582 in plain torch on c2w's device -- the sub-pixel modes the device kernel does not generate
    ('lefttop', 'random'; 'center' works too): rays, unit view directions taken BEFORE the NDC warp, optional warp."""
    rays_o, rays_d = get_rays(H, W, K, c2w, inverse_y=inverse_y, flip_x=flip_x, flip_y=flip_y, mode=mode)
    viewdirs = rays_d / rays_d.norm(dim=-1, keepdim=True)
    if ndc:
        rays_o, rays_d = ndc_rays(H, W, float(np.asarray(K)[0][0]), 1., rays_o, rays_d)
    return rays_o, rays_d, viewdirs


def _views(rgb_tr, train_poses, HW, Ks, ndc, inverse_y, flip_x, flip_y):
    from .dvgo import get_rays_of_a_view
    for img, c2w, (H, W), K in zip(rgb_tr, train_poses, HW, Ks):
        yield img, get_rays_of_a_view(H=H, W=W, K=K, c2w=c2w, ndc=ndc, inverse_y=inverse_y, flip_x=flip_x, flip_y=flip_y), int(H), int(W)


@torch.no_grad()
def get_training_rays(rgb_tr, train_poses, HW, Ks, ndc, inverse_y, flip_x, flip_y):
    """Per-view ray images of equally sized views: (rgb, rays_o, rays_d, viewdirs) each [N,H,W,3] + imsz (lib/dvgo.py:585-607)."""
    assert len(np.unique(HW, axis=0)) == 1 and len(np.unique(Ks.reshape(len(Ks), -1), axis=0)) == 1
    assert len(rgb_tr) == len(train_poses) == len(Ks) == len(HW)
    dev = rgb_tr.device
    cols = [[], [], []]
    for _, rays, _, _ in _views(rgb_tr, train_poses, HW, Ks, ndc, inverse_y, flip_x, flip_y):
        for c, r in zip(cols, rays):
            c.append(r.to(dev))
    return (rgb_tr, *[torch.stack(c) for c in cols], [1] * len(rgb_tr))


def _gather(rgb_tr_ori, train_poses, HW, Ks, ndc, inverse_y, flip_x, flip_y, select):
    assert len(rgb_tr_ori) == len(train_poses) == len(Ks) == len(HW)
    dev = rgb_tr_ori[0].device
    parts, imsz = [[], [], [], []], []
    for img, rays, H, W in _views(rgb_tr_ori, train_poses, HW, Ks, ndc, inverse_y, flip_x, flip_y):
        assert tuple(img.shape[:2]) == (H, W)
        m = select(rays)                                     # bool [H,W] or None (= all pixels)
        pick = (lambda t: t.to(dev)[m]) if m is not None else (lambda t: t.to(dev).flatten(0, 1))
        for p, t in zip(parts, (img, *rays)):
            p.append(pick(t))
        imsz.append(int(parts[0][-1].shape[0]))
    return (*[torch.cat(p) for p in parts], imsz)


@torch.no_grad()
def get_training_rays_flatten(rgb_tr_ori, train_poses, HW, Ks, ndc, inverse_y, flip_x, flip_y):
    """All pixels of all (differently sized) views as flat [N,3] lists + pixels per view (lib/dvgo.py:610-641)."""
    return _gather(rgb_tr_ori, train_poses, HW, Ks, ndc, inverse_y, flip_x, flip_y, lambda rays: None)


@torch.no_grad()
def get_training_rays_in_maskcache_sampling(rgb_tr_ori, train_poses, HW, Ks, ndc, inverse_y, flip_x, flip_y, model, render_kwargs):
    """Only the pixels whose rays hit the coarse geometry (lib/dvgo.py:644-684)."""
    dev = rgb_tr_ori[0].device

    def select(rays):
        o, d = rays[0], rays[1]
        return torch.cat([model.hit_coarse_geo(rays_o=o[i:i + 64], rays_d=d[i:i + 64], **render_kwargs) for i in range(0, o.shape[0], 64)]).to(dev)
    return _gather(rgb_tr_ori, train_poses, HW, Ks, ndc, inverse_y, flip_x, flip_y, select)


def batch_indices_generator(N, BS):
    """Endless stream of index batches over a fresh random permutation per epoch (lib/dvgo.py:687-697; the last, short
    batch of an epoch is dropped and a new permutation starts)."""
    while True:
        perm = torch.from_numpy(np.random.permutation(N)).long()
        if BS >= N:
            yield perm[:BS]
            continue
        for top in range(0, N - BS + 1, BS):
            yield perm[top:top + BS]

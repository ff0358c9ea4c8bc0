"""Seeded synthetic scenes and cameras of SURVEY.md section 8(d) -- TEST INFRASTRUCTURE ONLY.

There is no dataset or checkpoint in the container, so every parity test and bench leg renders
random-weight scenes of the shapes BASELINE.json names.  Shapes come from the reference configs:

 cfg1  configs/default.py:82-105  coarse DVGO: density + 3-ch colour grid, no MLP (64^3 here)
 cfgA  configs/default.py:107-119 fine DVGO:   160^3, k0 12 ch, rgbnet 39->128->128->3, direct
 cfgB  configs/llff/llff_default_lg.py:33-44 + fern_lg_joint_l1.py:24-32
                                  LLFF MPI:    [384,384,256], k0 9 ch, rgbnet 15->64->64->3

Density regimes: FOG (N(0,1) density, mask all true -- nearly every in-box sample survives both
thresholds) and SHELL (-10 everywhere, +6 on a spherical shell 3 voxels thick; occupancy mask =
maxpool3(alpha) > thres as update_occupancy_cache, lib/dvgo.py:224-233).
Seeds: 0 density, 1 k0, 2 rgbnet (nn.Linear default init, last bias 0 as lib/dvgo.py:124).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import pipeline


def _linear_init(out_f, in_f, gen):
    """nn.Linear.reset_parameters: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(in), 1/sqrt(in)) for W and b."""
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=gen) * 2 - 1) * bound
    return w, b


def fill_rgbnet(st, seed=2):
    if st['rgbnet'] is None:
        return
    gen = torch.Generator().manual_seed(seed)
    layers = []
    for (w, b) in st['rgbnet']:
        layers.append(_linear_init(w.shape[0], w.shape[1], gen))
    layers[-1] = (layers[-1][0], torch.zeros_like(layers[-1][1]))
    st['rgbnet'] = layers


def fill_grids(st, regime='fog', seed_density=0, seed_k0=1, k0_scale=1.0):
    ws = list(st['density'].shape[2:])
    g = torch.Generator().manual_seed(seed_k0)
    st['k0'] = torch.randn(st['k0'].shape, generator=g) * k0_scale
    if regime == 'fog':
        g = torch.Generator().manual_seed(seed_density)
        st['density'] = torch.randn(st['density'].shape, generator=g)
        if st['kind'] == 'dmpigo':
            # the MPI bias grid already gives alpha ~ 1/mpi_depth per sample; widen it a little
            st['density'] = st['density'] * 2.0
        mask = torch.ones(ws, dtype=torch.bool)
    elif regime == 'shell':
        X, Y, Z = ws
        ii, jj, kk = torch.meshgrid(torch.arange(X), torch.arange(Y), torch.arange(Z), indexing='ij')
        c = torch.tensor([(X - 1) / 2, (Y - 1) / 2, (Z - 1) / 2])
        # radius 0.5 * half-extent in voxel units (anisotropic grids: per-axis normalised)
        rr = torch.sqrt((((ii - c[0]) / ((X - 1) / 2)) ** 2 + ((jj - c[1]) / ((Y - 1) / 2)) ** 2 +
                         ((kk - c[2]) / ((Z - 1) / 2)) ** 2))
        half_vox = 1.5 / (min(X, Y, Z) / 2)           # 3 voxels thick along the finest axis
        shell = (rr - 0.5).abs() <= half_vox
        den = torch.full(ws, -10.0)
        den[shell] = 6.0
        if st['kind'] == 'dmpigo':
            den[shell] = 12.0                          # the MPI bias is about -5.5: keep the shell opaque
        st['density'] = den[None, None].contiguous()
        # occupancy: maxpool3(alpha) > fast_color_thres, lib/dvgo.py:230-233 / lib/dmpigo.py:221-224
        if st['kind'] == 'dvgo':
            shift = float(st['act_shift'])
            dens = st['density']
        else:
            shift = 0.0
            dens = st['density'] + st['act_shift_grid']
        interval = float(st['voxel_size_ratio'])
        alpha = 1 - torch.pow(1 + torch.exp(dens + shift), -interval)
        alpha = F.max_pool3d(alpha, kernel_size=3, padding=1, stride=1)[0, 0]
        mask = alpha > st['fast_color_thres']
    else:
        raise ValueError(regime)
    st['mask_cache'] = pipeline.mask_grid_state(mask, st['xyz_min'], st['xyz_max'])
    return st


def make_cfg1(res=64, regime='fog'):
    """BASELINE.json configs[0]: coarse-shape DVGO, density + 3-ch colour grid, no MLP."""
    st = pipeline.dvgo_state([-1, -1, -1], [1, 1, 1], num_voxels=res ** 3, num_voxels_base=res ** 3,
                             alpha_init=1e-6, fast_color_thres=1e-7, rgbnet_dim=0)
    fill_grids(st, regime)
    if regime == 'fog':
        st['density'] = st['density'] * 3 + 10.0    # alpha_init=1e-6 shifts by -13.8: lift the fog into view
    return st


def make_cfgA(res=160, regime='fog', rgbnet_direct=True, width=128, depth=3, k0_dim=12, viewbase_pe=4):
    """BASELINE.json configs[1]: fine-stage DVGO of configs/default.py:107-119."""
    st = pipeline.dvgo_state([-1, -1, -1], [1, 1, 1], num_voxels=res ** 3, num_voxels_base=res ** 3,
                             alpha_init=1e-2, fast_color_thres=1e-4, rgbnet_dim=k0_dim,
                             rgbnet_direct=rgbnet_direct, rgbnet_depth=depth, rgbnet_width=width,
                             viewbase_pe=viewbase_pe)
    fill_grids(st, regime)
    fill_rgbnet(st)
    return st


def make_cfgB(xy=384, depth=256, regime='fog', width=64, k0_dim=9, viewbase_pe=0, spatial_pe=0):
    """BASELINE.json configs[2]: LLFF multiplane model (lib/dmpigo), SURVEY.md section 8(d) box."""
    xyz_min = [-1.5, -1.67, -1.0]
    xyz_max = [1.5, 1.67, 1.0]
    # num_voxels chosen so that world_size[:2] comes out near `xy` on both axes' mean
    st = pipeline.dmpigo_state(xyz_min, xyz_max, num_voxels=xy * xy * depth, mpi_depth=depth,
                               fast_color_thres=1.0 / depth / 5, rgbnet_dim=k0_dim, rgbnet_depth=3,
                               rgbnet_width=width, viewbase_pe=viewbase_pe, spatial_pe=spatial_pe)
    fill_grids(st, regime)
    fill_rgbnet(st)
    return st


def make_cfgC(res=64, regime='fog', width=128, k0_dim=12, viewbase_pe=4, bg_len=0.2):
    """Unbounded inward-facing scene: DirectContractedVoxGO (lib/dcvgo.py), fine-stage shape of
    configs/default.py with the contracted background shell."""
    st = pipeline.dcvgo_state([-1, -1, -1], [1, 1, 1], num_voxels=res ** 3, num_voxels_base=res ** 3,
                              alpha_init=1e-2, fast_color_thres=1e-4, bg_len=bg_len, rgbnet_dim=k0_dim,
                              rgbnet_depth=3, rgbnet_width=width, viewbase_pe=viewbase_pe)
    st['kind'] = 'dvgo'          # fill_grids only distinguishes the MPI bias; contracted == bounded here
    fill_grids(st, regime)
    st['kind'] = 'dcvgo'
    fill_rgbnet(st)
    return st


# ---------------------------------------------------------------------------------------------
# cameras
# ---------------------------------------------------------------------------------------------
def pose_spherical(theta, phi, radius):
    """Blender-convention orbit pose (restates lib/load_blender.py:10-35)."""
    def trans_t(t):
        return torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, t], [0, 0, 0, 1]], dtype=torch.float32)

    def rot_phi(p):
        return torch.tensor([[1, 0, 0, 0], [0, np.cos(p), -np.sin(p), 0], [0, np.sin(p), np.cos(p), 0],
                             [0, 0, 0, 1]], dtype=torch.float32)

    def rot_theta(th):
        return torch.tensor([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0],
                             [0, 0, 0, 1]], dtype=torch.float32)
    c2w = trans_t(radius)
    c2w = rot_phi(phi / 180. * np.pi) @ c2w
    c2w = rot_theta(theta / 180. * np.pi) @ c2w
    c2w = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=torch.float32) @ c2w
    return c2w


def blender_camera(H, W, theta=30.0, phi=-30.0, radius=4.0):
    """Pinhole of the nerf_synthetic loader: focal 1111 px at 800 px width (lib/load_blender.py:77)."""
    focal = 1111.111 * (W / 800.0)
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], dtype=np.float32)
    return K, pose_spherical(theta, phi, radius)[:3, :4]


def blender_rays(H, W, theta=30.0, phi=-30.0, radius=4.0, crop=None):
    K, c2w = blender_camera(H, W, theta, phi, radius)
    ro, rd, vd = pipeline.get_rays_of_a_view(H, W, K, c2w, ndc=False, inverse_y=False, flip_x=False, flip_y=False)
    if crop is not None:
        y0, y1, x0, x1 = crop
        ro, rd, vd = ro[y0:y1, x0:x1], rd[y0:y1, x0:x1], vd[y0:y1, x0:x1]
    return ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3).contiguous(), vd.reshape(-1, 3).contiguous()


def llff_camera(H, W, shift=(0.0, 0.0, 0.0)):
    """Forward-facing camera looking down -z from near the origin (LLFF convention), focal ~ 0.8*W
    like the 4032x3024 fern capture (3260 px)."""
    focal = 0.8085 * W
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], dtype=np.float32)
    c2w = torch.tensor([[1, 0, 0, shift[0]], [0, 1, 0, shift[1]], [0, 0, 1, shift[2]]], dtype=torch.float32)
    return K, c2w


def llff_rays(H, W, shift=(0.05, -0.03, 0.0), crop=None):
    K, c2w = llff_camera(H, W, shift)
    ro, rd, vd = pipeline.get_rays_of_a_view(H, W, K, c2w, ndc=True, inverse_y=False, flip_x=False, flip_y=False)
    if crop is not None:
        y0, y1, x0, x1 = crop
        ro, rd, vd = ro[y0:y1, x0:x1], rd[y0:y1, x0:x1], vd[y0:y1, x0:x1]
    return ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3).contiguous(), vd.reshape(-1, 3).contiguous()


RENDER_KW_DVGO = dict(near=2.0, far=6.0, bg=1, stepsize=0.5, inverse_y=False, flip_x=False, flip_y=False,
                      render_depth=True)
RENDER_KW_DCVGO = dict(near=0.2, far=1e9, bg=1, stepsize=0.5, inverse_y=False, flip_x=False, flip_y=False,
                       render_depth=True)
RENDER_KW_MPI = dict(near=0, far=1, bg=0, stepsize=1.0, inverse_y=False, flip_x=False, flip_y=False,
                     render_depth=True)

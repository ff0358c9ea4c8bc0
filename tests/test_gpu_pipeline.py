"""GPU: the callers either side of the kernels -- render_viewpoints (both scripts' contracts) and the
full 4K-NeRF inference chain marcher -> VC-Decoder (run_sr.py:1344-1395) on a small synthetic scene,
against the oracle pipeline end to end."""
import numpy as np
import pytest
import torch

import k4nerf
from k4nerf import render
from oracle import ops, pipeline, scenes, sftnet
from helpers import make_state, model_from_state

pytestmark = pytest.mark.gpu


def _oracle_frame(st, H, W, K, c2w, ndc, kw):
    ro, rd, vd = pipeline.get_rays_of_a_view(H, W, K, c2w, ndc, kw['inverse_y'], kw['flip_x'], kw['flip_y'])
    r = pipeline.render_rays_chunked(st, ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3).contiguous(),
                                     vd.reshape(-1, 3).contiguous(), ops.CpuOps, chunk=8192, **kw)
    return {k: v.reshape(H, W, -1) for k, v in r.items()}


def test_render_viewpoints_contracts(cuda_device):
    st = make_state('cfgA', res=32, regime='fog')
    m = model_from_state(st, cuda_device)
    H, W = 24, 40
    poses, Ks = [], []
    for th in (30.0, 120.0):
        K, c2w = scenes.blender_camera(H, W, theta=th)
        poses.append(c2w.numpy()); Ks.append(K)
    HW = np.array([[H, W]] * 2)
    Ks = np.array(Ks)
    kw = dict(scenes.RENDER_KW_DVGO)
    rgbs, depths, bgmaps, psnrs, ssims, lp = render.render_viewpoints(m, np.array(poses), HW, Ks, False, kw)
    assert rgbs.shape == (2, H, W, 3) and depths.shape == (2, H, W, 1) and bgmaps.shape == (2, H, W, 1)
    assert rgbs.min() >= 0 and rgbs.max() <= 1 and psnrs == []
    out = render.render_viewpoints_sr(m, np.array(poses), HW, Ks, False, kw, gt_imgs=[rgbs[0], rgbs[1]])
    rgbs2, depths2, bgmaps2, psnrs2, viewdirs_all, feats = out
    assert feats.shape == (2, H, W, 3) and len(viewdirs_all) == 2 and viewdirs_all[0].shape == (H * W, 3)
    # the tcgen05 kernels accumulate a ray's samples with shared-memory atomics: run-to-run results
    # agree to fp32 rounding, not bit for bit
    assert np.abs(rgbs - rgbs2).max() < 1e-5 and all(p > 100 or np.isinf(p) for p in psnrs2)
    ref = _oracle_frame(st, H, W, Ks[0], torch.as_tensor(poses[0]), False, kw)
    assert pipeline.psnr(torch.from_numpy(feats[0]), ref['rgb_marched']) >= 70.0
    assert np.abs(depths[0] - ref['depth'].numpy()).max() < 1e-4
    # render_factor halves the resolution like run.py:83-87
    r4 = render.render_viewpoints(m, np.array(poses), HW, Ks, False, kw, render_factor=2)
    assert r4[0].shape == (2, H // 2, W // 2, 3)


def test_full_4k_nerf_chain_mpi_plus_decoder(cuda_device):
    """LLFF-style MPI render (ndc rays) -> rgb_feature + depth -> SFTNet.tile_process x4."""
    dev = cuda_device
    st = make_state('cfgB', xy=48, depth=32, regime='fog')
    m = model_from_state(st, dev)
    H, W = 24, 32
    K, c2w = scenes.llff_camera(H, W, (0.05, -0.03, 0.0))
    kw = dict(scenes.RENDER_KW_MPI)
    out = render.render_viewpoints_sr(m, np.array([c2w.numpy()]), np.array([[H, W]]), np.array([K]), True, kw)
    rgbs, depths, bgmaps, _, _, feats = out
    ref = _oracle_frame(st, H, W, K, c2w, True, kw)
    assert pipeline.psnr(torch.from_numpy(feats[0]), ref['rgb_marched']) >= 70.0
    sd = sftnet.random_state_dict(seed=3)
    net = k4nerf.SFTNet(3, 4, 64, 5, 32, 1)
    net.load_state_dict(sd)
    net = net.to(dev)
    # run_sr.py:1362-1367,1385: rgb_feature -> [1,3,H,W]; depth -> [1,H,W]
    x = torch.from_numpy(feats[0]).movedim(-1, 0).unsqueeze(0).to(dev)
    cond = torch.from_numpy(depths[0]).movedim(-1, 0).to(dev)
    sr = net.tile_process(x, cond, tile_size=16)
    assert sr.shape == (1, 3, 4 * H, 4 * W) and sr.device.type == 'cpu'
    x_ref = ref['rgb_marched'].movedim(-1, 0).unsqueeze(0)
    c_ref = ref['depth'].movedim(-1, 0)
    sr_ref = sftnet.tile_process(sd, x_ref, c_ref, 16)
    assert pipeline.psnr(sr, sr_ref) >= 60.0


def test_render_frame_4k_device_resident(cuda_device):
    """f-1: the device-resident chain equals render_viewpoints_sr + tile_process + clamp."""
    dev = cuda_device
    st = make_state('cfgA', res=32, regime='fog')
    m = model_from_state(st, dev)
    sd = sftnet.random_state_dict(seed=3, scale=1.0)
    net = k4nerf.SFTNet(3, 4, 64, 5, 32, 1)
    net.load_state_dict(sd)
    net = net.to(dev)
    H, W = 20, 28
    K, c2w = scenes.blender_camera(H, W)
    kw = dict(scenes.RENDER_KW_DVGO)
    sr, lr = render.render_frame_4k(m, net, H, W, K, c2w, False, kw, test_tile=16)
    assert sr.shape == (3, 4 * H, 4 * W) and sr.is_cuda and float(sr.min()) >= 0 and float(sr.max()) <= 1
    _, depths, _, _, _, feats = render.render_viewpoints_sr(m, np.array([c2w.numpy()]), np.array([[H, W]]), np.array([K]), False, kw)
    x = torch.from_numpy(feats[0]).movedim(-1, 0).unsqueeze(0).to(dev)
    cond = torch.from_numpy(depths[0]).movedim(-1, 0).to(dev)
    ref = net.tile_process(x, cond, tile_size=16).squeeze(0).clamp(0, 1)
    assert (sr.cpu() - ref).abs().max().item() < 2e-3
    u8, _ = render.render_frame_4k(m, net, H, W, K, c2w, False, kw, test_tile=16, out_u8=True)
    assert u8.dtype == torch.uint8 and u8.shape == (4 * H, 4 * W, 3)
    assert (u8.cpu().float() - (ref * 255).floor().permute(1, 2, 0)).abs().max().item() <= 1

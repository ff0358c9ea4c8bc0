"""Test helpers: turn an oracle scene state (oracle/pipeline.py dict) into a product model."""
import os

import torch

from oracle import pipeline, scenes

PRETRAINED_SR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_ref',
                             'RealESRNet_x4plus.pth')


def model_from_state(st, device=None):
    """Build the k4nerf module with the reference constructor kwargs and load the oracle's tensors.
    (k4nerf is imported here, not at module level: bench.py's reference arm uses the scene helpers of
    this file and must not map libk4nerf.so.)"""
    import k4nerf
    if st['kind'] == 'dvgo':
        nvox = int(st['density'].shape[2] * st['density'].shape[3] * st['density'].shape[4])
        kw = dict(xyz_min=st['xyz_min'].tolist(), xyz_max=st['xyz_max'].tolist(),
                  num_voxels=st['_num_voxels'], num_voxels_base=st['_num_voxels_base'],
                  alpha_init=st['_alpha_init'], fast_color_thres=st['fast_color_thres'],
                  rgbnet_dim=st['rgbnet_dim'], rgbnet_direct=st['rgbnet_direct'],
                  mask_cache_world_size=list(st['mask_cache']['mask'].shape))
        if st['rgbnet'] is not None:
            kw.update(rgbnet_depth=len(st['rgbnet']), rgbnet_width=st['rgbnet'][0][0].shape[0],
                      viewbase_pe=len(st['viewfreq']))
        m = k4nerf.DirectVoxGO(**kw)
        del nvox
    elif st['kind'] == 'dcvgo':
        kw = dict(xyz_min=st['_fg_min'], xyz_max=st['_fg_max'], num_voxels=st['_num_voxels'],
                  num_voxels_base=st['_num_voxels_base'], alpha_init=st['_alpha_init'],
                  fast_color_thres=st['fast_color_thres'], bg_len=st['bg_len'], rgbnet_dim=st['rgbnet_dim'],
                  mask_cache_world_size=list(st['mask_cache']['mask'].shape))
        if st['rgbnet'] is not None:
            kw.update(rgbnet_depth=len(st['rgbnet']), rgbnet_width=st['rgbnet'][0][0].shape[0],
                      viewbase_pe=len(st['viewfreq']))
        m = k4nerf.DirectContractedVoxGO(**kw)
    else:
        kw = dict(xyz_min=st['xyz_min'].tolist(), xyz_max=st['xyz_max'].tolist(),
                  num_voxels=st['_num_voxels'], mpi_depth=st['mpi_depth'],
                  fast_color_thres=st['fast_color_thres'], rgbnet_dim=st['rgbnet_dim'],
                  mask_cache_world_size=list(st['mask_cache']['mask'].shape),
                  act_type='relu', mode_type='mlp')
        if st['rgbnet'] is not None:
            kw.update(rgbnet_depth=len(st['rgbnet']), rgbnet_width=st['rgbnet'][0][0].shape[0],
                      viewbase_pe=len(st['viewfreq']), spatial_pe=len(st['posfreq']))
        m = k4nerf.DirectMPIGO(**kw)
    sd = m.state_dict()
    sd['density.grid'] = st['density'].clone()
    sd['k0.grid'] = st['k0'].clone()
    sd['mask_cache.mask'] = st['mask_cache']['mask'].clone()
    if st['rgbnet'] is not None:
        lin_keys = sorted({k.rsplit('.', 1)[0] for k in sd if k.startswith('rgbnet.')},
                          key=lambda s: [int(x) for x in s.split('.')[1:]])
        assert len(lin_keys) == len(st['rgbnet'])
        for k, (w, b) in zip(lin_keys, st['rgbnet']):
            sd[k + '.weight'] = w.clone()
            sd[k + '.bias'] = b.clone()
    m.load_state_dict(sd)
    if device is not None:
        m = m.to(device)
    return m


def make_state(name, **kw):
    """Named scene constructors that also remember the reference constructor kwargs."""
    if name == 'cfg1':
        res = kw.get('res', 32)
        st = scenes.make_cfg1(res=res, regime=kw.get('regime', 'fog'))
        st.update(_num_voxels=res ** 3, _num_voxels_base=res ** 3, _alpha_init=1e-6)
    elif name == 'cfgA':
        res = kw.pop('res', 48)
        st = scenes.make_cfgA(res=res, **kw)
        st.update(_num_voxels=res ** 3, _num_voxels_base=res ** 3, _alpha_init=1e-2)
    elif name == 'cfgC':
        res = kw.pop('res', 48)
        st = scenes.make_cfgC(res=res, **kw)
        st.update(_num_voxels=res ** 3, _num_voxels_base=res ** 3, _alpha_init=1e-2, _fg_min=[-1, -1, -1], _fg_max=[1, 1, 1])
    elif name == 'cfgB':
        xy, depth = kw.pop('xy', 48), kw.pop('depth', 32)
        st = scenes.make_cfgB(xy=xy, depth=depth, **kw)
        st.update(_num_voxels=xy * xy * depth)
    else:
        raise ValueError(name)
    return st


def rays_for(st, H, W, **kw):
    if st['kind'] == 'dcvgo':
        return scenes.blender_rays(H, W, **kw), dict(scenes.RENDER_KW_DCVGO)
    if st['kind'] == 'dvgo':
        return scenes.blender_rays(H, W, **kw), dict(scenes.RENDER_KW_DVGO)
    return scenes.llff_rays(H, W, **kw), dict(scenes.RENDER_KW_MPI)


def compare(ours, ref, n_rays):
    """Parity numbers between the fused kernel's outputs and the oracle's."""
    out = {}
    for k in ('rgb_marched', 'alphainv_last', 'depth'):
        if k in ref and k in ours:
            a, b = ours[k].detach().cpu().double().reshape(n_rays, -1), ref[k].detach().cpu().double().reshape(n_rays, -1)
            out[k + '_maxabs'] = (a - b).abs().max().item() if a.numel() else 0.0
            out[k + '_psnr'] = pipeline.psnr(a, b) if a.numel() else float('inf')
    return out


def pretrained_sr_state_dict(seed=3):
    """SFTNet parameters as run_sr.py:663 builds them: a freshly constructed net (here the seeded
    random_state_dict, so every consumer rebuilds the same set) with pretrained/RealESRNet_x4plus.pth
    loaded on top, strict=False (the RRDB convs / conv_first / conv_body / up / hr / last come from the
    checkpoint, the SFT layers and CondNet keep their initial values).  Needs oracle/_ref/RealESRNet_x4plus.pth
    (copied by oracle/build_ref.py where /root/reference exists); returns None when it is absent."""
    from oracle import sftnet
    if not os.path.exists(PRETRAINED_SR):
        return None
    sd = sftnet.random_state_dict(seed=seed)
    ck = torch.load(PRETRAINED_SR, map_location='cpu', weights_only=False)
    ck = ck['params_ema'] if 'params_ema' in ck else ck['params']
    n = 0
    for k, v in ck.items():
        k = k[7:] if k.startswith('module.') else k
        if k in sd and sd[k].shape == v.shape:
            sd[k] = v.float().clone()
            n += 1
    assert n >= 160, n
    return sd


def ref_forward_chunked(st_dev, ro, rd, vd, kw, ref_ops, chunk=8192):
    """The reference's render_viewpoints chunk loop (run_sr.py:121-124: 8192-ray chunks, concatenated)
    over oracle.pipeline.forward on an op backend; returns the concatenated per-ray outputs, the per-ray
    visited-sample counts and the summed sample statistics."""
    outs = {'rgb_marched': [], 'alphainv_last': [], 'depth': [], '_ray_stats': []}
    stats = {}
    for s in range(0, ro.shape[0], chunk):
        r = pipeline.forward(st_dev, ro[s:s + chunk].contiguous(), rd[s:s + chunk].contiguous(),
                             vd[s:s + chunk].contiguous(), ref_ops, stats=stats, **kw)
        for k in outs:
            if k in r:
                outs[k].append(r[k].clone())
    return {k: torch.cat(v) for k, v in outs.items() if v}, stats

"""Training-side forward of the three models: the reference's materialised
pipeline -- flat sample lists, three compactions, autograd through density / k0 / rgbnet -- built on the
op-level kernels (render_utils_cuda drop-in + autograd_ops) so the reference's training loop
(`scene_rep_reconstruction*`, run_sr.py) can run on this library.  ATen supplies grid_sample (and its
gradient scatter into the grids) and the nn.Linear layers, exactly as in the reference.

Returns the reference's key set: ``alphainv_last, weights, rgb_marched, rgb_feature, raw_alpha,
raw_rgb, ray_id`` (+ ``n_max, s`` for DirectMPIGO; + ``n_max, s, t, step_id, raw_density, wsum_mid`` for
DirectContractedVoxGO, lib/dcvgo.py:358-371 -- run_sr.py:531-532 reads ``t`` / ``raw_density`` when
weight_nearclip > 0; + ``depth`` on request).
Inference never comes here: `forward` under no_grad uses the fused marcher."""
import torch

from . import _lib, render_utils_cuda as ops
from .autograd_ops import Raw2Alpha, Alphas2Weights


def _take(mask, *tensors):
    return tuple(t[mask] for t in tensors)


def _segment_sum(src, ray_id, n_rays):
    """torch_scatter.segment_coo(reduce='sum') over a sorted index == index_add into zeros."""
    out = torch.zeros((n_rays,) + tuple(src.shape[1:]), device=src.device, dtype=src.dtype)
    return out.index_add_(0, ray_id, src)


def _view_embedding(viewdirs, viewfreq):
    e = (viewdirs.unsqueeze(-1) * viewfreq).flatten(-2)
    return torch.cat([viewdirs, e.sin(), e.cos()], -1)


# ---- samplers: world points, their ray / step ids, and the per-sample depth coordinate ------------
def _sample_dvgo(m, rays_o, rays_d, kw):
    stepdist = kw['stepsize'] * m.voxel_size                                   # lib/dvgo.py:310
    n_max = int((m.max_world_size - 1) / kw['stepsize']) + 1
    pts, outside, ray_id, step_id, _, _, _ = ops.sample_pts_on_rays(
        rays_o.contiguous(), rays_d.contiguous(), m.xyz_min, m.xyz_max, kw['near'], 1e9, stepdist)
    pts, ray_id, step_id = _take(~outside, pts, ray_id, step_id)
    return pts, ray_id, step_id, n_max, (step_id + 0.5) / n_max, {}


def _sample_mpi(m, rays_o, rays_d, kw):
    assert kw['near'] == 0 and kw['far'] == 1                                   # lib/dmpigo.py:275
    n_max = int((m.mpi_depth - 1) / kw['stepsize']) + 1
    pts, outside = ops.sample_ndc_pts_on_rays(rays_o.contiguous(), rays_d.contiguous(), m.xyz_min, m.xyz_max, n_max)
    inside = ~outside
    n = inside.shape[0]
    ray_id = torch.arange(n, device=pts.device).view(-1, 1).expand_as(inside)[inside]
    step_id = torch.arange(n_max, device=pts.device).view(1, -1).expand_as(inside)[inside]
    return pts[inside], ray_id, step_id, n_max, (step_id + 0.5) / n_max, {}


def _sample_contracted(m, rays_o, rays_d, kw):
    dev = rays_o.device
    o = (rays_o - m.scene_center) / m.scene_radius                              # lib/dcvgo.py:237-262
    d = rays_d / rays_d.norm(dim=-1, keepdim=True)
    t = m.sample_t(kw['stepsize'], dev)
    pts = o[:, None, :] + d[:, None, :] * t[None, :, None]
    norm = pts.abs().amax(dim=-1, keepdim=True)
    inner = norm <= 1
    pts = torch.where(inner, pts, pts / norm * ((1 + m.bg_len) - m.bg_len / norm))
    keep = inner.squeeze(-1).clone()
    thres = (2 + 2 * m.bg_len) / m.world_len * kw['stepsize'] * 0.95           # lib/dcvgo.py:283-285
    dist = (pts[:, 1:] - pts[:, :-1]).norm(dim=-1)
    keep[:, 1:] |= ops.cumdist_thres(dist.contiguous(), thres)
    n, n_max = keep.shape
    ray_id = torch.arange(n, device=dev).view(-1, 1).expand(n, n_max)[keep]
    step_id = torch.arange(n_max, device=dev).view(1, -1).expand(n, n_max)[keep]
    tt = t[None].expand(n, n_max)[keep]
    return pts[keep], ray_id, step_id, n_max, 1 - 1 / (1 + tt), {'t': tt, 'inner': inner.squeeze(-1)[keep]}


_SAMPLERS = {_lib.K4_KIND_DVGO: _sample_dvgo, _lib.K4_KIND_DMPIGO: _sample_mpi, _lib.K4_KIND_DCVGO: _sample_contracted}


def forward_samples(m, rays_o, rays_d, viewdirs, global_step=None, is_train=None, **kw):
    """The reference forward (lib/dvgo.py:327-448, lib/dmpigo.py:292-427, lib/dcvgo.py:264-382) with
    autograd; works under no_grad too (then it is the un-fused restatement of the marcher)."""
    assert rays_o.dim() == 2 and rays_o.shape[-1] == 3, 'Only suuport point queries in [N, 3] format'
    kind = m._k4_kind
    n_rays = rays_o.shape[0]
    pts, ray_id, step_id, n_max, s, aux = _SAMPLERS[kind](m, rays_o, rays_d, kw)
    interval = float(kw['stepsize'] * m.voxel_size_ratio)

    def carry(sel):               # per-sample extras that follow every compaction (DCVGO: t, inner_mask, density)
        for k in aux:
            aux[k] = aux[k][sel]

    occupied = m.mask_cache(pts)
    pts, ray_id, step_id, s = _take(occupied, pts, ray_id, step_id, s)
    carry(occupied)

    density = m.density(pts)
    if kind == _lib.K4_KIND_DCVGO:
        aux['density'] = density
    if kind == _lib.K4_KIND_DMPIGO:
        alpha = Raw2Alpha.apply((density + m.act_shift(pts)).flatten(), 0, interval)
    else:
        alpha = Raw2Alpha.apply(density.flatten(), float(m.act_shift), interval)
    thres = m.fast_color_thres
    if thres > 0:
        sel = alpha > thres
        pts, ray_id, step_id, s, alpha = _take(sel, pts, ray_id, step_id, s, alpha)
        carry(sel)

    weights, alphainv_last = Alphas2Weights.apply(alpha, ray_id, n_rays)
    if thres > 0:
        sel = weights > thres
        pts, ray_id, step_id, s, alpha, weights = _take(sel, pts, ray_id, step_id, s, alpha, weights)
        carry(sel)

    k0 = m.k0(pts)
    if m.rgbnet is None:
        rgb = torch.sigmoid(k0)
    else:
        vemb = _view_embedding(viewdirs, m.viewfreq)[ray_id]
        if kind == _lib.K4_KIND_DMPIGO:
            p = ((pts - m.xyz_min) / (m.xyz_max - m.xyz_min)).flip((-1,)) * 2 - 1
            pe = (p.unsqueeze(-1) * m.posfreq).flatten(-2)
            rgb = torch.sigmoid(m.rgbnet(torch.cat([k0, p, pe.sin(), pe.cos(), vemb], -1)))
        elif getattr(m, 'rgbnet_direct', True):
            rgb = torch.sigmoid(m.rgbnet(torch.cat([k0, vemb], -1)))
        else:
            rgb = torch.sigmoid(m.rgbnet(torch.cat([k0[:, 3:], vemb], -1)) + k0[:, :3])

    rgb_feature = _segment_sum(weights.unsqueeze(-1) * rgb, ray_id, n_rays)
    training = is_train if (kind == _lib.K4_KIND_DCVGO and is_train is not None) else (global_step is not None)
    if kw.get('rand_bkgd', False) and training and kind != _lib.K4_KIND_DVGO:
        rgb_marched = rgb_feature + alphainv_last.unsqueeze(-1) * torch.rand_like(rgb_feature)
    else:
        rgb_marched = rgb_feature
        rgb_marched += alphainv_last.unsqueeze(-1) * kw['bg']          # in place: rgb_feature aliases it (reference quirk)
    ret = {'alphainv_last': alphainv_last, 'weights': weights, 'rgb_marched': rgb_marched, 'rgb_feature': rgb_feature,
           'raw_alpha': alpha, 'raw_rgb': rgb, 'ray_id': ray_id}
    if kind != _lib.K4_KIND_DVGO:
        ret.update(n_max=n_max, s=s)
    if kind == _lib.K4_KIND_DCVGO:          # lib/dcvgo.py:352-371
        inner = aux['inner']
        ret.update(t=aux['t'], step_id=step_id, raw_density=aux['density'],
                   wsum_mid=_segment_sum(weights[inner], ray_id[inner], n_rays))
    if kw.get('render_depth', False):
        with torch.no_grad():
            ret['depth'] = _segment_sum(weights * s, ray_id, n_rays)
    return ret

"""Multi-GPU rendering of one frame: row-band sharding + one NCCL all-gather.

The reference has no multi-GPU path (its only NCCL code is vendored and unreachable, SURVEY.md
section 2.3), so this is new: rays are independent and the scene is read-only, hence every rank
holds a full replica of the scene (217 MB for 160^3 / 12 ch), marches its share of the image rows
with the same fused kernel -- 8-row blocks (the kernel's tile height) dealt round-robin over the
ranks, so that every rank gets the same mix of long and short rays -- and the shares are exchanged
with ONE ``all_gather_into_tensor`` of a packed ``[rgb(3) | depth | alphainv]`` buffer over
NVLink/NVSwitch.  No reduction crosses
GPUs, so results are bit-identical to the single-GPU render of the same rows.

Works with any ``torch.distributed`` backend (``nccl`` on GPUs; ``gloo`` in the CPU tests, where a
stand-in render function is injected because the kernel itself needs a GPU).
"""
import torch
import torch.distributed as dist


ROW_BLOCK = 8      # height of the marcher's warpgroup tile (16x8 pixels): the unit of row interleaving


def cyclic_rows(H, rank, world_size, block=ROW_BLOCK):
    """Row indices of `rank` under block-cyclic sharding: 8-row blocks b with b % world == rank.
    Contiguous bands give the middle ranks the long rays (load imbalance = max over ranks); cyclic
    blocks give every rank the same mix.  Within a block the rows stay adjacent, so the kernel's
    16x8 pixel tiles keep their spatial locality."""
    nblk = (H + block - 1) // block
    rows = [torch.arange(b * block, min((b + 1) * block, H)) for b in range(rank, nblk, world_size)]
    return torch.cat(rows) if rows else torch.zeros(0, dtype=torch.long)


def cyclic_pad_rows(H, world_size, block=ROW_BLOCK):
    nblk = (H + block - 1) // block
    return ((nblk + world_size - 1) // world_size) * block


def unpack_frame_cyclic(gathered, H, W, world_size, block=ROW_BLOCK):
    """[world, 5 * n_pad] (block-cyclic bands) -> dict of full-frame tensors in image order."""
    rows_pad = cyclic_pad_rows(H, world_size, block)
    n_pad = rows_pad * W
    g = gathered.view(world_size, 5 * n_pad)
    dev = gathered.device
    rgb = torch.empty(H, W, 3, device=dev, dtype=gathered.dtype)
    depth = torch.empty(H, W, device=dev, dtype=gathered.dtype)
    ainv = torch.empty(H, W, device=dev, dtype=gathered.dtype)
    for r in range(world_size):
        rows = cyclic_rows(H, r, world_size, block).to(dev)
        k = rows.numel()
        rgb[rows] = g[r, :3 * n_pad].view(rows_pad, W, 3)[:k]
        depth[rows] = g[r, 3 * n_pad:4 * n_pad].view(rows_pad, W)[:k]
        ainv[rows] = g[r, 4 * n_pad:5 * n_pad].view(rows_pad, W)[:k]
    return {'rgb_marched': rgb.reshape(H * W, 3), 'depth': depth.reshape(H * W), 'alphainv_last': ainv.reshape(H * W)}


def band_rows(H, world_size):
    """Rows per band: every rank gets the same (padded) number of rows so that one
    fixed-size all-gather suffices."""
    return (H + world_size - 1) // world_size


def band_range(H, rank, world_size):
    rows = band_rows(H, world_size)
    r0 = min(rank * rows, H)
    r1 = min(r0 + rows, H)
    return r0, r1


def pack_band(out, n_band_rays, n_pad_rays):
    """[5 * n_pad] buffer: rgb (3n) | depth (n) | alphainv (n); the tail of a short band is zero."""
    dev = out['rgb_marched'].device
    buf = torch.zeros(5 * n_pad_rays, device=dev, dtype=torch.float32)
    buf[0:3 * n_band_rays] = out['rgb_marched'].reshape(-1)
    if 'depth' in out:
        buf[3 * n_pad_rays:3 * n_pad_rays + n_band_rays] = out['depth']
    buf[4 * n_pad_rays:4 * n_pad_rays + n_band_rays] = out['alphainv_last']
    return buf


def unpack_frame(gathered, H, W, world_size):
    """[world, 5 * n_pad] -> dict of full-frame tensors (padding rows dropped)."""
    rows = band_rows(H, world_size)
    n_pad = rows * W
    g = gathered.view(world_size, 5 * n_pad)
    rgb = g[:, :3 * n_pad].reshape(world_size * rows, W, 3)[:H]
    depth = g[:, 3 * n_pad:4 * n_pad].reshape(world_size * rows, W)[:H]
    ainv = g[:, 4 * n_pad:5 * n_pad].reshape(world_size * rows, W)[:H]
    return {'rgb_marched': rgb.reshape(H * W, 3), 'depth': depth.reshape(H * W), 'alphainv_last': ainv.reshape(H * W)}


def render_frame_sharded(render_fn, rays_o, rays_d, viewdirs, H, W, group=None, gather=True, layout='cyclic'):
    """Render this rank's rows with ``render_fn(ro, rd, vd, image_hw) -> dict`` and all-gather the
    packed bands.  ``rays_*`` are the FULL frame ``[H*W, 3]`` (replicated or generated per rank); only
    this rank's rows are read.  ``layout``: 'cyclic' (8-row blocks dealt round-robin, load balanced,
    default) or 'bands' (contiguous).  Returns the full-frame dict on every rank (``gather=True``) or
    this rank's band dict."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if layout == 'cyclic' and world > 1:
        rows = cyclic_rows(H, rank, world).to(rays_o.device)
        k = rows.numel()
        sel = lambda t: t.view(H, W, 3)[rows].reshape(-1, 3)
        if k > 0:
            out = render_fn(sel(rays_o), sel(rays_d), sel(viewdirs), (k, W))
        else:
            dev = rays_o.device
            out = {'rgb_marched': torch.zeros(0, 3, device=dev), 'depth': torch.zeros(0, device=dev),
                   'alphainv_last': torch.zeros(0, device=dev)}
        if not gather:
            return out
        n_pad = cyclic_pad_rows(H, world) * W
        buf = pack_band(out, k * W, n_pad)
        gathered = torch.empty(world * 5 * n_pad, device=buf.device, dtype=torch.float32)
        dist.all_gather_into_tensor(gathered, buf, group=group)
        return unpack_frame_cyclic(gathered, H, W, world)
    r0, r1 = band_range(H, rank, world)
    sl = slice(r0 * W, r1 * W)
    n_band = (r1 - r0) * W
    if n_band > 0:
        out = render_fn(rays_o[sl], rays_d[sl], viewdirs[sl], (r1 - r0, W))
    else:
        dev = rays_o.device
        out = {'rgb_marched': torch.zeros(0, 3, device=dev), 'depth': torch.zeros(0, device=dev),
               'alphainv_last': torch.zeros(0, device=dev)}
    if not gather or world == 1:
        return out
    n_pad = band_rows(H, world) * W
    buf = pack_band(out, n_band, n_pad)
    gathered = torch.empty(world * 5 * n_pad, device=buf.device, dtype=torch.float32)
    dist.all_gather_into_tensor(gathered, buf, group=group)
    return unpack_frame(gathered, H, W, world)

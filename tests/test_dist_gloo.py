"""CPU: the N>1 path (row-band sharding + all-gather) with world_size 2 on the gloo backend.
The fused kernel needs a GPU, so a deterministic stand-in render function is injected; what is
tested is the sharding / packing / gathering logic of k4nerf/dist.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_render(ro, rd, vd, image_hw):
    # any per-ray function: results must land at the right place of the gathered frame
    return {'rgb_marched': torch.stack([ro[:, 0] + rd[:, 1], ro[:, 1] * 2, vd[:, 2] - 1], -1),
            'depth': ro[:, 2] * 0.5, 'alphainv_last': rd[:, 0] + 3}


def _worker(rank, world, port, H, W, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, '4k-nerf_b200'))
    from k4nerf import dist as kdist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    ro, rd, vd = (torch.randn(H * W, 3, generator=g) for _ in range(3))
    ref = _fake_render(ro, rd, vd, (H, W))
    ok = True
    for layout in ('cyclic', 'bands'):
        full = kdist.render_frame_sharded(_fake_render, ro, rd, vd, H, W, layout=layout)
        ok = ok and all(torch.equal(full[k], ref[k]) for k in ('rgb_marched', 'depth', 'alphainv_last'))
    # the device-resident driver's path: only the rank's rows are "generated", the render writes into the
    # packed buffer's views, one all-gather, image-order transpose
    frame = kdist.CyclicFrame(H, W, torch.device('cpu'))
    make = lambda rows: tuple(t.view(H, W, 3)[rows.long()].reshape(-1, 3) for t in (ro, rd, vd))

    def render_into(a, b, c, hw, out):
        r = _fake_render(a, b, c, hw)
        for k in out:
            out[k].copy_(r[k])
    for _ in range(2):                                   # buffers are reused across frames
        full = frame.render(make, render_into)
        ok = ok and all(torch.equal(full[k], ref[k]) for k in ('rgb_marched', 'depth', 'alphainv_last'))
    r0, r1 = kdist.band_range(H, rank, world)
    q.put((rank, ok, r0, r1))
    dist.destroy_process_group()


@pytest.mark.parametrize('H,W', [(8, 6), (7, 5), (1, 9), (37, 4)])
def test_row_band_sharding_world2_gloo(H, W):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, H, W, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res
    rows = sorted((r0, r1) for _, _, r0, r1 in res)
    assert rows[0][0] == 0 and rows[-1][1] == H and rows[0][1] == rows[1][0]


def test_band_arithmetic():
    import sys
    from k4nerf import dist as kdist
    for H in (1, 2, 7, 756, 3024):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                r0, r1 = kdist.band_range(H, r, world)
                assert 0 <= r0 <= r1 <= H and r1 - r0 <= kdist.band_rows(H, world)
                cover += list(range(r0, r1))
            assert cover == list(range(H))
            cyc = sorted(int(x) for r in range(world) for x in kdist.cyclic_rows(H, r, world))
            assert cyc == list(range(H))
            assert all(kdist.cyclic_rows(H, r, world).numel() <= kdist.cyclic_pad_rows(H, world) for r in range(world))


# ---------------------------------------------------------------------------------------------
# decoder sharding (k4nerf.dist.sr_units / sr_decode_sharded)
# ---------------------------------------------------------------------------------------------
def _fake_decoder(x, c):
    """Stand-in x4 'decoder' with a 2-pixel receptive radius and zero padding (5x5 box filter of
    img + cond, then nearest x4): tiling, halo and zero-padding behaviour are all exercised."""
    import torch.nn.functional as F
    y = x + c
    k = torch.ones(3, 1, 5, 5, dtype=x.dtype) / 25
    y = F.conv2d(y, k, padding=2, groups=3)
    return F.interpolate(y, scale_factor=4, mode='nearest')


def _tile_process_local(fn, img, cond, tile, pad):
    import math
    _, C, H, W = img.shape
    out = img.new_zeros((1, C, 4 * H, 4 * W))
    c4 = cond.unsqueeze(0)
    for ty in range(math.ceil(H / tile)):
        for tx in range(math.ceil(W / tile)):
            x0, y0 = tx * tile, ty * tile
            x1, y1 = min(x0 + tile, W), min(y0 + tile, H)
            x0p, x1p, y0p, y1p = max(x0 - pad, 0), min(x1 + pad, W), max(y0 - pad, 0), min(y1 + pad, H)
            o = fn(img[:, :, y0p:y1p, x0p:x1p], c4[:, :, y0p:y1p, x0p:x1p])
            out[:, :, 4 * y0:4 * y1, 4 * x0:4 * x1] = o[:, :, 4 * (y0 - y0p):4 * (y0 - y0p + y1 - y0), 4 * (x0 - x0p):4 * (x0 - x0p + x1 - x0)]
    return out


def _sr_worker(rank, world, port, H, W, tile, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, '4k-nerf_b200'))
    from k4nerf import dist as kdist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(9)
    img = torch.randn(1, 3, H, W, generator=g)
    cond = torch.randn(1, H, W, generator=g)
    ref = _tile_process_local(_fake_decoder, img, cond, tile, 1)
    full = kdist.sr_decode_sharded(_fake_decoder, img, cond, tile, tile_pad=1, scale=4, halo=2)
    q.put((rank, bool(torch.equal(full, ref))))
    dist.destroy_process_group()


@pytest.mark.parametrize('H,W,tile', [(23, 31, 16), (40, 12, 64)])
def test_decoder_sharding_world2_gloo(H, W, tile):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sr_worker, args=(r, 2, port, H, W, tile, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_sr_units_cover_and_halo():
    """Units tile the frame exactly once for every world size; splitting a tile into row parts with
    halo >= receptive radius reproduces the un-split tile (single process, simulated ranks)."""
    from k4nerf import dist as kdist
    for (H, W, tile) in ((756, 1008, 510), (100, 70, 510), (33, 65, 32)):
        for world in (1, 2, 3, 4, 8, 16):
            units = kdist.sr_units(H, W, tile, 10, world, 80)
            cover = torch.zeros(H, W, dtype=torch.int32)
            for u in units:
                y0, y1, x0, x1 = u['dst']
                cover[y0:y1, x0:x1] += 1
                sa, sb, xa, xb = u['src']
                assert 0 <= sa <= y0 and y1 <= sb <= H and 0 <= xa <= x0 and x1 <= xb <= W
                assert u['keep'] == (y0 - sa, x0 - xa)
            assert bool((cover == 1).all()), (H, W, tile, world)
            assert len(units) >= min(world, len(kdist.sr_units(H, W, tile, 10, 1, 80)))
    g = torch.Generator().manual_seed(2)
    img, cond = torch.randn(1, 3, 41, 37, generator=g), torch.randn(1, 41, 37, generator=g)
    ref = _tile_process_local(_fake_decoder, img, cond, 24, 1)
    for world in (1, 3, 8):
        units = kdist.sr_units(41, 37, 24, 1, world, 2)
        out = torch.zeros_like(ref)
        for u in units:
            sa, sb, xa, xb = u['src']
            o = _fake_decoder(img[:, :, sa:sb, xa:xb], cond.unsqueeze(0)[:, :, sa:sb, xa:xb])
            ky, kx = u['keep']
            y0, y1, x0, x1 = u['dst']
            out[:, :, 4 * y0:4 * y1, 4 * x0:4 * x1] = o[:, :, 4 * ky:4 * (ky + y1 - y0), 4 * kx:4 * (kx + x1 - x0)]
        assert torch.equal(out, ref), world


def test_sr_plan_is_the_cached_units_and_assignment():
    from k4nerf import dist as kdist
    for world in (1, 2, 3, 4, 8):
        units, assign = kdist.sr_plan(756, 1008, 510, 10, world, 80)
        assert units == kdist.sr_units(756, 1008, 510, 10, world, 80)
        assert assign == kdist.sr_assign(units, world)
        again = kdist.sr_plan(756, 1008, 510, 10, world, 80)
        assert again[0] is units and again[1] is assign                  # no search on the per-frame path


def test_peer_buffers_fall_back_without_cuda_and_frame_target_layout():
    """PeerBuffers.create is collective and answers None off the GPUs of one node (here: no process group, CPU) -- the
    callers then take the all-gather path the gloo tests above exercise; FrameTarget fills k4_frame_dst."""
    from k4nerf import dist as kdist
    assert kdist.PeerBuffers.create(1024, 'cpu') is None
    t = kdist.FrameTarget([0x1000, 0x2000, 0x3000], rank=1, world=3, W=72, n_full=3 * 24 * 72)
    d = t.frame_dst
    assert (d.n_dst, d.rank, d.world, d.frame_w, d.n_full) == (3, 1, 3, 72, 3 * 24 * 72)
    assert [d.d_frame[i] for i in range(3)] == [0x1000, 0x2000, 0x3000] and d.d_frame[3] is None
    H, W, world = 50, 72, 3
    n_full = world * kdist.cyclic_pad_rows(H, world) * W
    full = torch.arange(5 * n_full, dtype=torch.float32)
    v = kdist.frame_views(full, H, W, n_full)
    assert v['rgb_marched'].shape == (H * W, 3) and v['rgb_marched'][1, 0] == 3.0
    assert v['depth'][0] == 3 * n_full and v['alphainv_last'][H * W - 1] == 4 * n_full + H * W - 1
    # the kernel's row map (k4_store_ray): local row r of rank q is image row ((r // 8) * world + q) * 8 + r % 8
    for q in range(world):
        rows = kdist.cyclic_rows(H, q, world).tolist()
        assert rows == [((r // 8) * world + q) * 8 + r % 8 for r in range(len(rows))]

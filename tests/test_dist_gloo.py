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

"""CPU, property-based (hypothesis): the sharding arithmetic of k4nerf/dist.py for arbitrary frame sizes and world sizes --
block-cyclic rows partition the image, the packed-band <-> image-order maps are inverse to each other and agree with the
row map the marcher's frame stores use (k4_store_ray), decoder units tile the frame exactly once and every rank's load
respects the plan's bound."""
import torch
from hypothesis import given, settings, strategies as st

from k4nerf import dist as kdist


@settings(max_examples=60, deadline=None)
@given(H=st.integers(1, 700), W=st.integers(1, 40), world=st.integers(1, 8))
def test_cyclic_rows_partition_and_image_order_roundtrip(H, W, world):
    rows = [kdist.cyclic_rows(H, r, world) for r in range(world)]
    assert sorted(int(x) for t in rows for x in t) == list(range(H))
    rows_pad = kdist.cyclic_pad_rows(H, world)
    assert rows_pad % kdist.ROW_BLOCK == 0 and all(t.numel() <= rows_pad for t in rows)
    n_pad, n_full = rows_pad * W, world * rows_pad * W
    # every rank packs f(image row, column) for its rows; gather + unpack must put it at the image position
    val = lambda y, x, c: (y * W + x) * 5 + c
    bands = []
    for r in range(world):
        k = rows[r].numel()
        out = {'rgb_marched': torch.tensor([[val(int(y), x, c) for c in range(3)] for y in rows[r] for x in range(W)],
                                           dtype=torch.float32).view(k * W, 3),
               'depth': torch.tensor([val(int(y), x, 3) for y in rows[r] for x in range(W)], dtype=torch.float32),
               'alphainv_last': torch.tensor([val(int(y), x, 4) for y in rows[r] for x in range(W)], dtype=torch.float32)}
        bands.append(kdist.pack_band(out, k * W, n_pad))
    full_buf = torch.full((5 * n_full,), -1.0)
    got = kdist.unpack_frame_cyclic(torch.cat(bands), H, W, world, out=full_buf)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing='ij')
    base = ((yy * W + xx) * 5).reshape(-1).float()
    assert torch.equal(got['rgb_marched'], torch.stack([base, base + 1, base + 2], -1))
    assert torch.equal(got['depth'], base + 3) and torch.equal(got['alphainv_last'], base + 4)
    # the views of an image-order frame (peer mode) address the same elements the unpack wrote
    v = kdist.frame_views(full_buf, H, W, n_full)
    assert all(torch.equal(v[k], got[k]) for k in got)
    # the kernel's row map: local row r of rank q -> image row ((r // 8) * world + q) * 8 + r % 8
    for q in range(world):
        assert rows[q].tolist() == [((r // 8) * world + q) * 8 + r % 8 for r in range(rows[q].numel())]


@settings(max_examples=40, deadline=None)
@given(H=st.integers(1, 300), W=st.integers(1, 300), tile=st.integers(8, 128), pad=st.integers(0, 6),
       world=st.integers(1, 9), halo=st.integers(0, 20))
def test_decoder_units_tile_the_frame_once_and_the_plan_is_consistent(H, W, tile, pad, world, halo):
    units, assign = kdist.sr_plan(H, W, tile, pad, world, halo)
    cover = torch.zeros(H, W, dtype=torch.int32)
    for u in units:
        y0, y1, x0, x1 = u['dst']
        sa, sb, xa, xb = u['src']
        ky, kx = u['keep']
        assert 0 <= sa <= y0 < y1 <= sb <= H and 0 <= xa <= x0 < x1 <= xb <= W      # the source window holds the kept block
        assert sa + ky == y0 and xa + kx == x0                                        # ... at offset `keep`
        cover[y0:y1, x0:x1] += 1
    assert bool((cover == 1).all())
    assert sorted(i for a in assign for i in a) == list(range(len(units))) and len(assign) == world

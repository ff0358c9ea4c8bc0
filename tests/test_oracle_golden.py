"""CPU: the oracle against the committed golden vectors (tests/golden/, made by make_golden.py)."""
import os

import pytest
import torch

from oracle import ops, pipeline, sftnet
from helpers import make_state, rays_for

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _load(name):
    return torch.load(os.path.join(GOLD, name), map_location='cpu', weights_only=False)


def _sftnet_inputs():
    g = torch.Generator().manual_seed(11)
    x = torch.rand(1, 3, 20, 28, generator=g) * 1.2 - 0.1
    c = torch.rand(1, 1, 20, 28, generator=g)
    xt = torch.rand(1, 3, 24, 40, generator=g) * 1.2 - 0.1
    ct = torch.rand(1, 24, 40, generator=g)
    return x, c, xt, ct


def test_sftnet_oracle_matches_reference_module_output():
    """sftnet_ref.pt was produced by the REFERENCE's lib/sr_esrnet.py SFTNet: this pins the
    functional restatement (same ATen conv kernels => tolerance is rounding-order only)."""
    gold = _load('sftnet_ref.pt')
    sd = sftnet.random_state_dict(seed=gold['param_seed'])
    x, c, xt, ct = _sftnet_inputs()
    y = sftnet.sftnet_forward(sd, x, c)
    assert y.shape == gold['forward'].shape
    assert (y - gold['forward']).abs().max().item() <= 2e-5
    yt = sftnet.tile_process(sd, xt, ct, tile_size=16, tile_pad=10)
    assert yt.shape == gold['tile_process'].shape
    assert (yt - gold['tile_process']).abs().max().item() <= 2e-5


def test_tile_plan_geometry():
    # 1008x756 with tile 510 / pad 10 -> 2x2 tiles
    plan = sftnet.tile_plan(756, 1008, 510, 10)
    assert len(plan) == 4
    assert plan[0][:4] == (0, 520, 0, 520) and plan[3][:4] == (500, 756, 500, 1008)
    assert sum((p[1] - p[0]) * (p[3] - p[2]) for p in plan) == 797728      # padded LR pixels, SURVEY 8(d)


CASES = {
    'cfgA_fog': ('cfgA', dict(res=24, regime='fog'), (16, 20)),
    'cfgA_shell': ('cfgA', dict(res=24, regime='shell'), (16, 20)),
    'cfgB_fog': ('cfgB', dict(xy=24, depth=16, regime='fog'), (12, 16)),
    'cfg1_fog': ('cfg1', dict(res=16), (16, 16)),
}


@pytest.mark.parametrize('name', sorted(CASES))
def test_marcher_oracle_matches_golden(name):
    kind, kw, hw = CASES[name]
    gold = _load(f'marcher_{name}.pt')
    st = make_state(kind, **kw)
    (ro, rd, vd), rkw = rays_for(st, *hw)
    stats = {}
    r = pipeline.forward(st, ro, rd, vd, ops.CpuOps, stats=stats, **rkw)
    assert stats == gold['stats']
    assert torch.allclose(r['rgb_marched'], gold['rgb_marched'], atol=2e-6, rtol=0)
    assert torch.allclose(r['depth'], gold['depth'], atol=2e-6, rtol=0)
    assert torch.allclose(r['alphainv_last'], gold['alphainv_last'], atol=2e-6, rtol=0)
    if 'N_steps' in gold:
        assert torch.equal(r['_N_steps'].to(torch.int32), gold['N_steps'])


def test_chunked_equals_whole(tmp_path):
    """The reference renders 8192-ray chunks (run_sr.py:121-124); chunking must not change a ray."""
    st = make_state('cfgA', res=16, regime='fog')
    (ro, rd, vd), kw = rays_for(st, 12, 12)
    whole = pipeline.forward(st, ro, rd, vd, ops.CpuOps, **kw)
    parts = pipeline.render_rays_chunked(st, ro, rd, vd, ops.CpuOps, chunk=50, **kw)
    assert torch.equal(whole['rgb_marched'], parts['rgb_marched'])
    assert torch.equal(whole['alphainv_last'], parts['alphainv_last'])


def test_c_ops_against_double_precision_formulas():
    """The C restatement vs straightforward float64 evaluation of the same formulas."""
    g = torch.Generator().manual_seed(0)
    ro = torch.randn(500, 3, generator=g) * 0.2 + torch.tensor([0., 0., 3.])
    rd = torch.randn(500, 3, generator=g) * 0.2 + torch.tensor([0., 0., -1.])
    mn, mx = torch.tensor([-1., -1., -1.]), torch.tensor([1., 1., 1.])
    t_min, t_max = ops.CpuOps.infer_t_minmax(ro, rd, mn, mx, 0.2, 1e9)
    a = (mx.double() - ro.double()) / rd.double()
    b = (mn.double() - ro.double()) / rd.double()
    tm = torch.minimum(a, b).amax(-1).clamp(0.2, 1e9)
    tM = torch.maximum(a, b).amin(-1).clamp(0.2, 1e9)
    assert torch.allclose(t_min.double(), tm, rtol=1e-5, atol=1e-5)
    assert torch.allclose(t_max.double(), tM, rtol=1e-5, atol=1e-5)
    d = torch.linspace(-12, 12, 1001)
    e, al = ops.CpuOps.raw2alpha(d, -4.595, 0.5)
    ref = 1 - (1 + torch.exp(d.double() - 4.595)) ** -0.5
    assert torch.allclose(al.double(), ref, atol=2e-7)
    # alpha2weight: transmittance prefix product with the T < 1e-3 early-out
    alpha = torch.full((40,), 0.3)
    rid = torch.zeros(40, dtype=torch.int64)
    w, T, last, i_s, i_e = ops.CpuOps.alpha2weight(alpha, rid, 1)
    k = int(i_e[0])
    assert 0 < k < 40 and float(last[0]) < 1e-3 and float(T[k - 1]) >= 1e-3
    assert torch.all(w[k:] == 0) and torch.all(T[k:] == 1)
    assert abs(float(w.sum() + last[0]) - 1.0) < 1e-6


def test_cumdist_thres_restatement():
    """k4o_cumdist_thres vs a literal loop over lib/cuda/ub360_utils_kernel.cu:12-32."""
    import numpy as np
    g = torch.Generator().manual_seed(11)
    dist = torch.rand(7, 29, generator=g) * 0.03
    thres = 0.05
    got = ops.CpuOps.cumdist_thres(dist, thres)
    want = torch.zeros_like(got)
    for r in range(dist.shape[0]):
        cum = np.float32(0)
        for i in range(dist.shape[1]):
            cum = np.float32(cum + np.float32(dist[r, i].item()))
            over = bool(cum > np.float32(thres))
            want[r, i] = over
            if over:
                cum = np.float32(0)
    assert torch.equal(got, want) and int(got.sum()) > 0


def test_dcvgo_oracle_runs_and_is_deterministic():
    from helpers import make_state, rays_for
    st = make_state('cfgC', res=24, regime='shell')
    (ro, rd, vd), kw = rays_for(st, 12, 16, radius=0.6)
    s1, s2 = {}, {}
    a = pipeline.forward(st, ro, rd, vd, ops.CpuOps, stats=s1, **kw)
    b = pipeline.forward(st, ro, rd, vd, ops.CpuOps, stats=s2, **kw)
    assert torch.equal(a['rgb_marched'], b['rgb_marched']) and s1 == s2
    assert 0 < s1['S_c'] <= s1['S_d'] <= s1['S_m'] and a['depth'].min() >= 0 and a['depth'].max() <= 1

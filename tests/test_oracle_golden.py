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


def test_sftnet_oracle_with_pretrained_weights_matches_reference_module():
    """sftnet_pretrained_ref.pt: the REFERENCE module with pretrained/RealESRNet_x4plus.pth loaded
    strict=False as run_sr.py:663 does.  The checkpoint travels as oracle/_ref/RealESRNet_x4plus.pth
    (git-ignored, copied by oracle/build_ref.py); skipped where it is absent."""
    import pytest
    from helpers import pretrained_sr_state_dict
    sd = pretrained_sr_state_dict()
    if sd is None:
        pytest.skip('oracle/_ref/RealESRNet_x4plus.pth missing')
    gold = _load('sftnet_pretrained_ref.pt')
    g = torch.Generator().manual_seed(gold['input_seed'])
    x = torch.rand(1, 3, 40, 48, generator=g) * 1.2 - 0.1
    c = torch.rand(1, 1, 40, 48, generator=g)
    y = sftnet.sftnet_forward(sd, x, c)
    assert y.shape == gold['forward'].shape
    assert (y - gold['forward']).abs().max().item() <= 2e-5


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
    # DirectContractedVoxGO (f-3): camera inside the inner cube, both regimes
    'cfgC_fog': ('cfgC', dict(res=24, regime='fog'), (12, 16), dict(radius=0.6)),
    'cfgC_shell': ('cfgC', dict(res=24, regime='shell'), (12, 16), dict(radius=0.6)),
}


@pytest.mark.parametrize('name', sorted(CASES))
def test_marcher_oracle_matches_golden(name):
    kind, kw, hw = CASES[name][:3]
    gold = _load(f'marcher_{name}.pt')
    st = make_state(kind, **kw)
    (ro, rd, vd), rkw = rays_for(st, *hw, **(CASES[name][3] if len(CASES[name]) > 3 else {}))
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


def test_total_variation_oracle_vs_numpy():
    """k4o_total_variation_add_grad against an independent numpy statement of the clamped-difference TV
    gradient (lib/cuda/total_variation_kernel.cu:13-38)."""
    import numpy as np
    g = torch.Generator().manual_seed(2)
    p = (torch.randn(1, 2, 5, 4, 6, generator=g) * 1.5)
    grad0 = torch.randn(p.shape, generator=g)
    grad0[torch.rand(p.shape, generator=g) < 0.5] = 0
    for dense in (True, False):
        got = grad0.clone()
        ops.CpuOps.total_variation_add_grad(p.contiguous(), got, 0.6, 1.2, 0.3, dense)
        P = p.numpy().astype(np.float32)
        want = grad0.numpy().copy()
        w = [np.float32(0.3) / np.float32(6), np.float32(1.2) / np.float32(6), np.float32(0.6) / np.float32(6)]   # axis 2 (i): wz ... axis 4 (k): wx
        w = {2: np.float32(0.3 / 1) / np.float32(6), 3: np.float32(1.2) / np.float32(6), 4: np.float32(0.6) / np.float32(6)}
        add = np.zeros_like(P)
        for ax in (4, 3, 2):                                   # the kernel adds k, then j, then i terms
            for sgn in (-1, +1):                               # minus neighbour first, then plus neighbour
                nb = np.roll(P, -sgn, axis=ax)
                term = w[ax] * np.clip(P - nb, -1, 1)
                idx = [slice(None)] * 5
                idx[ax] = 0 if sgn == -1 else P.shape[ax] - 1
                term[tuple(idx)] = 0
                add = (add + term).astype(np.float32)
        want2 = want + add
        if not dense:
            want2 = np.where(want == 0, want, want2)
        assert np.array_equal(got.numpy(), want2.astype(np.float32)), dense


def test_adam_oracle_vs_numpy():
    """k4o_adam_upd (three variants) against numpy float32 arithmetic with the reference's expression shapes
    (lib/cuda/adam_upd_kernel.cu:19-23,76)."""
    import numpy as np
    f = np.float32
    g = torch.Generator().manual_seed(4)
    n = 257
    p0, m0, v0 = torch.randn(n, generator=g), torch.randn(n, generator=g) * 0.1, (torch.randn(n, generator=g) * 0.1) ** 2
    grad = torch.randn(n, generator=g)
    grad[torch.rand(n, generator=g) < 0.6] = 0
    perlr = torch.rand(n, generator=g)
    b1, b2, lr, eps, step = f(0.9), f(0.99), f(0.1), f(1e-8), 3
    ss = f(lr * f(np.sqrt(f(1 - f(np.power(b2, f(step)))))) / f(1 - f(np.power(b1, f(step)))))
    G = grad.numpy()
    m = np.array([f(np.float64(b1) * np.float64(a) + np.float64(f(f(1) - b1) * gg)) for a, gg in zip(m0.numpy(), G)], dtype=np.float32)   # fma
    v = np.array([f(np.float64(b2) * np.float64(a) + np.float64(f(f(f(1) - b2) * gg) * gg)) for a, gg in zip(v0.numpy(), G)], dtype=np.float32)
    for variant in ('plain', 'masked', 'perlr'):
        p, mm, vv = p0.clone(), m0.clone(), v0.clone()
        if variant == 'plain':
            ops.CpuOps.adam_upd(p, grad, mm, vv, step, float(b1), float(b2), float(lr), float(eps))
        elif variant == 'masked':
            ops.CpuOps.masked_adam_upd(p, grad, mm, vv, step, float(b1), float(b2), float(lr), float(eps))
        else:
            ops.CpuOps.adam_upd_with_perlr(p, grad, mm, vv, perlr, step, float(b1), float(b2), float(lr), float(eps))
        num = (f(ss) * perlr.numpy()) * m if variant == 'perlr' else f(ss) * m
        pw = (p0.numpy() - (num / (np.sqrt(v) + eps)).astype(np.float32)).astype(np.float32)
        touched = (G != 0) if variant == 'masked' else np.ones(n, bool)
        assert np.array_equal(mm.numpy(), np.where(touched, m, m0.numpy())), variant
        assert np.array_equal(vv.numpy(), np.where(touched, v, v0.numpy())), variant
        assert np.array_equal(p.numpy(), np.where(touched, pw, p0.numpy())), variant

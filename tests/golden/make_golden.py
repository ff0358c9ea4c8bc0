"""Generate the committed golden fixtures (run in the BUILD CONTAINER, where /root/reference exists).

    python tests/golden/make_golden.py

 sftnet_ref.pt     output of the REFERENCE's own lib/sr_esrnet.py SFTNet (imported from
                   /root/reference, CPU, fp32) on seeded inputs and seeded parameters
                   (oracle.sftnet.random_state_dict(seed=3), loaded with load_state_dict(strict=True)):
                   one plain forward on a 20x28 tile and one tile_process (tile 16, pad 10) on 24x40.
                   This pins oracle/sftnet.py to the reference.
 marcher_*.pt      outputs of the marcher oracle (oracle/pipeline.py + CpuOps) on small seeded
                   scenes: rgb_marched / depth / alphainv_last / per-ray step counts and the sample
                   counters S_m, S_d, S_c.  The reference's marcher cannot run in the container (CUDA only),
                   so these pin the oracle against ITSELF across refactors; the oracle's arithmetic is
                   pinned to the reference's compiled kernels on the GPU box (tests/test_gpu_ref_ops.py).
Only outputs + seeds are stored (a few hundred KB); inputs are regenerated from the seeds.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, '4k-nerf_b200'))

from oracle import ops, pipeline, sftnet  # noqa: E402


def sftnet_inputs():
    g = torch.Generator().manual_seed(11)
    x = torch.rand(1, 3, 20, 28, generator=g) * 1.2 - 0.1      # rgb_feature is unclamped
    c = torch.rand(1, 1, 20, 28, generator=g)
    xt = torch.rand(1, 3, 24, 40, generator=g) * 1.2 - 0.1
    ct = torch.rand(1, 24, 40, generator=g)
    return x, c, xt, ct


def make_sftnet_golden():
    sys.path.insert(0, '/root/reference')
    from lib import sr_esrnet                       # the reference's own module
    torch.manual_seed(0)
    net = sr_esrnet.SFTNet(n_in_colors=3, scale=4, num_feat=64, num_block=5, num_grow_ch=32, num_cond=1, dswise=False)
    sd = sftnet.random_state_dict(seed=3)
    net.load_state_dict(sd, strict=True)
    net.eval()
    x, c, xt, ct = sftnet_inputs()
    with torch.no_grad():
        y = net(x, c)
        import io, contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            yt = net.tile_process(xt, ct, tile_size=16, tile_pad=10)
    torch.save({'forward': y.clone(), 'tile_process': yt.clone(), 'param_seed': 3, 'input_seed': 11,
                'source': 'reference lib/sr_esrnet.py SFTNet, torch %s CPU' % torch.__version__},
               os.path.join(HERE, 'sftnet_ref.pt'))
    print('sftnet_ref.pt', tuple(y.shape), tuple(yt.shape))


def sftnet_pretrained_inputs():
    g = torch.Generator().manual_seed(12)
    x = torch.rand(1, 3, 40, 48, generator=g) * 1.2 - 0.1
    c = torch.rand(1, 1, 40, 48, generator=g)
    return x, c


def make_sftnet_pretrained_golden():
    """The REFERENCE module with the shipped pretrained/RealESRNet_x4plus.pth loaded exactly as
    run_sr.py:663 does (load_network(strict=False) on top of the constructor's parameters; here the
    non-pretrained SFT / CondNet parameters are the seeded set so that a consumer can rebuild them)."""
    sys.path.insert(0, '/root/reference')
    from lib import sr_esrnet
    import io, contextlib
    net = sr_esrnet.SFTNet(n_in_colors=3, scale=4, num_feat=64, num_block=5, num_grow_ch=32, num_cond=1, dswise=False)
    net.load_state_dict(sftnet.random_state_dict(seed=3), strict=True)
    with contextlib.redirect_stdout(io.StringIO()):
        net.load_network(load_path='/root/reference/pretrained/RealESRNet_x4plus.pth', device='cpu', strict=False)
    net.eval()
    x, c = sftnet_pretrained_inputs()
    with torch.no_grad():
        y = net(x, c)
    torch.save({'forward': y.clone(), 'param_seed': 3, 'input_seed': 12,
                'source': 'reference lib/sr_esrnet.py SFTNet + pretrained/RealESRNet_x4plus.pth (strict=False), torch %s CPU' % torch.__version__},
               os.path.join(HERE, 'sftnet_pretrained_ref.pt'))
    print('sftnet_pretrained_ref.pt', tuple(y.shape), float(y.min()), float(y.max()))


MARCHER_CASES = {
    'cfgA_fog': ('cfgA', dict(res=24, regime='fog'), (16, 20)),
    'cfgA_shell': ('cfgA', dict(res=24, regime='shell'), (16, 20)),
    'cfgB_fog': ('cfgB', dict(xy=24, depth=16, regime='fog'), (12, 16)),
    'cfg1_fog': ('cfg1', dict(res=16), (16, 16)),
    # DirectContractedVoxGO (f-3): camera inside the inner cube, both regimes
    'cfgC_fog': ('cfgC', dict(res=24, regime='fog'), (12, 16), dict(radius=0.6)),
    'cfgC_shell': ('cfgC', dict(res=24, regime='shell'), (12, 16), dict(radius=0.6)),
}


def make_marcher_golden():
    from helpers import make_state, rays_for
    for name, case in MARCHER_CASES.items():
        kind, kw, hw = case[:3]
        st = make_state(kind, **kw)
        (ro, rd, vd), rkw = rays_for(st, *hw, **(case[3] if len(case) > 3 else {}))
        stats = {}
        r = pipeline.forward(st, ro, rd, vd, ops.CpuOps, stats=stats, **rkw)
        out = {'rgb_marched': r['rgb_marched'].clone(), 'depth': r['depth'].clone(),
               'alphainv_last': r['alphainv_last'].clone(), 'stats': dict(stats)}
        if '_N_steps' in r:
            out['N_steps'] = r['_N_steps'].to(torch.int32)
        torch.save(out, os.path.join(HERE, f'marcher_{name}.pt'))
        print(name, stats)


if __name__ == '__main__':
    make_sftnet_golden()
    make_sftnet_pretrained_golden()
    make_marcher_golden()

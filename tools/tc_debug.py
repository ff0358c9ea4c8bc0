"""Development probe for the tcgen05 marcher: run one configuration under a watchdog."""
import os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, '4k-nerf_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
from helpers import make_state, model_from_state, rays_for

def main():
    res, H, W, regime, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
    linear = len(sys.argv) > 6 and sys.argv[6] == 'linear'
    dev = torch.device('cuda', 0)
    st = make_state('cfgA', res=res, regime=regime)
    (ro, rd, vd), kw = rays_for(st, H, W)
    m = model_from_state(st, dev)
    ro, rd, vd = ro.to(dev), rd.to(dev), vd.to(dev)
    ref = m.render_rays(ro, rd, vd, kw, image_hw=None if linear else (H, W), mlp_mode='f16', debug=True)
    torch.cuda.synchronize()
    t = time.time()
    out = m.render_rays(ro, rd, vd, kw, image_hw=None if linear else (H, W), mlp_mode=mode, debug=True)
    torch.cuda.synchronize()
    dt = time.time() - t
    print(json.dumps({'res': res, 'hw': [H, W], 'regime': regime, 'mode': mode, 'dbg': os.environ.get('K4_TC_DBG', '0'),
                      's': round(dt, 4), 'counters': out['counters'].cpu().tolist(), 'ref_counters': ref['counters'].cpu().tolist(),
                      'maxabs_vs_f16': (out['rgb_marched'] - ref['rgb_marched']).abs().max().item()}), flush=True)

main()

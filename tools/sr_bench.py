"""VC-Decoder timing: k4nerf.SFTNet.tile_process (tcgen05 convs) vs the same network run through
torch/cuDNN (the reference's execution path: oracle.sftnet on CUDA tensors), 1008x756 -> 4032x3024,
tile 510 / pad 10 (run_sr.py --test_tile 510)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, '4k-nerf_b200')):
    sys.path.insert(0, p)
import k4nerf  # noqa: E402
from oracle import pipeline, sftnet  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    no_ref = '--no-ref' in sys.argv
    H, W = (int(args[0]), int(args[1])) if len(args) > 1 else (756, 1008)
    iters = 5
    sd = sftnet.random_state_dict(seed=3, scale=1.0)
    net = k4nerf.SFTNet(3, 4, 64, 5, 32, 1)
    net.load_state_dict(sd)
    net = net.to(dev)
    g = torch.Generator().manual_seed(1)
    img = torch.rand(1, 3, H, W, generator=g).to(dev)
    cond = torch.rand(1, H, W, generator=g).to(dev)
    flop = 2 * 5188864 * sum((p[1] - p[0]) * (p[3] - p[2]) for p in sftnet.tile_plan(H, W, 510, 10))

    def timeit(fn):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out = fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters, out

    ms, out = timeit(lambda: net.tile_process(img, cond, 510, to_cpu=False))
    rec = {'what': 'k4nerf SFTNet.tile_process (tcgen05), tiles on 2 streams', 'hw': [H, W], 'ms_per_frame': ms, 'tflops': flop / ms / 1e9,
           'pdl': os.environ.get('K4_SR_PDL', '1')}
    print(json.dumps(rec), flush=True)
    ms1, out1 = timeit(lambda: net.tile_process(img, cond, 510, to_cpu=False, streams=1))
    print(json.dumps({'what': 'same, tiles one after another on one stream', 'ms_per_frame': ms1, 'tflops': flop / ms1 / 1e9,
                      'identical_to_2_streams': bool(torch.equal(out, out1))}), flush=True)
    ms4, out4 = timeit(lambda: net.tile_process(img, cond, 510, to_cpu=False, streams=4))
    print(json.dumps({'what': 'same, all four tiles concurrently (4 streams)', 'ms_per_frame': ms4, 'tflops': flop / ms4 / 1e9,
                      'identical_to_2_streams': bool(torch.equal(out, out4))}), flush=True)
    if '--unit' in sys.argv:           # the decoder unit of one rank of an 8-GPU frame: half a 520x520 tile + halo
        x = img[:, :, :345, :520].contiguous(); c = cond.unsqueeze(0)[:, :, :345, :520].contiguous()
        o = torch.empty(3, 255 * 4, 510 * 4, device=dev)
        msu, _ = timeit(lambda: net.forward_roi(x, c, (0, 255, 0, 510), o))
        full = torch.empty(3, 345 * 4, 520 * 4, device=dev)
        msf, _ = timeit(lambda: net.forward_roi(x, c, (0, 345, 0, 520), full))
        print(json.dumps({'what': '345x520 unit: kept block 255x510 with per-layer row windows vs all rows', 'ms_roi': msu, 'ms_all_rows': msf,
                          'identical': bool(torch.equal(o, full[:, :1020, :2040]))}), flush=True)
    if no_ref:
        return
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    refs = {}
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        ms_ref, ref = timeit(lambda: sftnet.tile_process(sd_dev, img, cond, 510))
        refs[tf32] = ref
        p = pipeline.psnr(out.cpu(), ref)
        print(json.dumps({'what': f'torch/cuDNN same network, allow_tf32={tf32} (the reference path; includes its per-tile D2H)',
                          'ms_per_frame': ms_ref, 'tflops': flop / ms_ref / 1e9, 'psnr_k4_vs_this': p,
                          'out_absmax': ref.abs().max().item()}), flush=True)
    print(json.dumps({'what': 'cuDNN TF32 (reference default) vs cuDNN fp32: the arithmetic noise of the reference itself',
                      'psnr': pipeline.psnr(refs[True], refs[False])}), flush=True)


if __name__ == '__main__':
    main()

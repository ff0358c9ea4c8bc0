"""torchrun check (N GPUs of one node): the sharded full-frame driver returns, on every rank, exactly
the frame the single-GPU driver returns.  torchrun --nproc-per-node N tools/frame_sharded_check.py [--quick]
(K4_PEER=0 in the environment: the all-gather exchange instead of in-kernel peer stores; --quick: no timing loops)"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, '4k-nerf_b200'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import k4nerf  # noqa: E402
from k4nerf import render as krender  # noqa: E402
from k4nerf import dist as krender_dist  # noqa: E402
from helpers import make_state, model_from_state  # noqa: E402
from oracle import scenes, sftnet  # noqa: E402


def main():
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    H, W = 756, 1008
    quick = '--quick' in sys.argv
    res = {}
    order = ('fog', 'shell') if '--fog-first' in sys.argv else ('shell', 'fog')
    for regime in order:
        st = make_state('cfgA', res=160, regime=regime)
        model = model_from_state(st, dev)
        net = k4nerf.SFTNet(3, 4, 64, 5, 32, 1)
        net.load_state_dict(sftnet.random_state_dict(seed=3, scale=1.0))
        net = net.to(dev)
        kw = dict(scenes.RENDER_KW_DVGO)
        K, c2w = scenes.blender_camera(H, W)
        model.mlp_mode = 'f16x3'          # deterministic accumulation order per ray (the ws kernel's atomics are not)
        ref, lr_ref = krender.render_frame_4k(model, net, H, W, K, c2w, False, kw, test_tile=510)
        sr, lr = krender.render_frame_4k_sharded(model, net, H, W, K, c2w, False, kw, test_tile=510)
        same_lr = all(torch.equal(lr[k], lr_ref[k]) for k in ('rgb_marched', 'depth', 'alphainv_last'))
        same = bool(torch.equal(sr, ref))
        t = torch.tensor([int(same), int(same_lr)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        res[regime] = {'sr_identical_all_ranks': bool(t[0].item()), 'lr_identical_all_ranks': bool(t[1].item()),
                       'maxabs': (sr - ref).abs().max().item()}
        # a second and a third frame (the two peer frame buffers alternate): still the same pixels
        for _ in range(2):
            sr2, lr2 = krender.render_frame_4k_sharded(model, net, H, W, K, c2w, False, kw, test_tile=510)
            t2 = torch.tensor([int(torch.equal(sr2, ref)), int(torch.equal(lr2['rgb_marched'], lr_ref['rgb_marched']))], device=dev)
            if world > 1:
                dist.all_reduce(t2, op=dist.ReduceOp.MIN)
            res[regime]['sr_identical_all_ranks'] &= bool(t2[0].item())
            res[regime]['lr_identical_all_ranks'] &= bool(t2[1].item())
        frame = list(model.__dict__['_k4_cyclic_frames'].values())[0]
        res['exchange'] = {'marcher': 'peer_stores' if frame.peers is not None else 'all_gather',
                           'decoder': 'peer_stores' if any(v is not None for v in net.__dict__.get('_k4_peer_frames', {}).values()) else 'all_gather',
                           'peer_error': krender_dist.PeerBuffers.last_error}
        model.mlp_mode = 'auto'
        if quick:
            continue
        for _ in range(2):
            krender.render_frame_4k_sharded(model, net, H, W, K, c2w, False, kw, test_tile=510)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import time
        e0.record()
        h0 = time.perf_counter()
        for _ in range(5):
            krender.render_frame_4k_sharded(model, net, H, W, K, c2w, False, kw, test_tile=510)
        host_ms = (time.perf_counter() - h0) * 1e3 / 5          # time to ENQUEUE a frame (nothing waits for the GPU here)
        e1.record()
        torch.cuda.synchronize()
        tm = torch.tensor([e0.elapsed_time(e1) / 5, host_ms], device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        res[regime]['ms_per_frame'] = tm[0].item()
        res[regime]['host_enqueue_ms_per_frame'] = round(tm[1].item(), 3)
        # phase breakdown (max over ranks): the marcher + its all-gather, and the decoder + its all-gather, timed apart
        x = lr['rgb_marched'].view(H, W, 3).permute(2, 0, 1).unsqueeze(0).contiguous()
        cond = lr['depth'].view(1, H, W).contiguous()
        from k4nerf import dvgo as kdvgo
        kw2 = dict(kw); kw2['render_depth'] = True
        make = lambda rows: tuple(t.view(-1, 3) for t in kdvgo.get_rays_of_a_view(H, W, K, c2w, False, False, False, False, rows=rows, device=dev))
        fn = lambda ro, rd, vd, hw, out: model.render_rays(ro, rd, vd, kw2, image_hw=hw, out=out)
        phases = {'march_gather': lambda: frame.render(make, fn),
                  'march_only': lambda: (frame.k > 0) and fn(*make(frame.rows), (frame.k, W), frame.target()),
                  'decode_gather': lambda: net.tile_process_sharded(x, cond, tile_size=510)}
        if world > 1 and '--peer-breakdown' in sys.argv and net.__dict__.get('_k4_peer_frames'):
            # where does a peer-mode decode spend its time: the unit with / without the stores into the other ranks' frames,
            # with / without the barrier (timing only: without the stores the other ranks' frames are incomplete)
            pb = krender_dist.PeerBuffers
            real_sync, real_units = pb.sync, net.run_units
            dec = phases['decode_gather']
            def no_sync(): pb.sync = lambda self: None
            def with_sync(): pb.sync = real_sync
            def local_only(): net.run_units = lambda jobs, streams=2: real_units([j[:4] for j in jobs], streams)
            def with_peers(): net.__dict__.pop('run_units', None)
            variants = {'decode_peer_sync': (with_peers, with_sync), 'decode_peer_nosync': (with_peers, no_sync),
                        'decode_local_nosync': (local_only, no_sync), 'decode_local_sync': (local_only, with_sync)}
            for name, (a, b) in variants.items():
                a(); b()
                dec(); torch.cuda.synchronize(); dist.barrier()
                e0.record()
                for _ in range(5):
                    dec()
                e1.record()
                torch.cuda.synchronize()
                mine = torch.tensor([e0.elapsed_time(e1) / 5], device=dev)
                allr = [torch.zeros_like(mine) for _ in range(world)]
                dist.all_gather(allr, mine)
                res[regime][name + '_ms_by_rank'] = [round(t.item(), 3) for t in allr]
            with_peers(); with_sync()
        for name, f in phases.items():
            f(); torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0.record()
            for _ in range(5):
                f()
            e1.record()
            torch.cuda.synchronize()
            tp = torch.tensor([e0.elapsed_time(e1) / 5], device=dev)
            if world > 1:
                dist.all_reduce(tp, op=dist.ReduceOp.MAX)
            res[regime][name + '_ms'] = round(tp.item(), 3)
    if rank == 0:
        print(json.dumps({'n_gpus': world, 'frame': '1008x756 -> 4032x3024', **res}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

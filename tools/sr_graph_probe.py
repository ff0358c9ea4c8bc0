"""Probe: how much of the decoder's per-tile time is launch gaps?  Times SFTNet.forward on one 520x520
tile eagerly and as a replayed CUDA graph (torch.cuda.CUDAGraph around the same C-ABI call)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, '4k-nerf_b200'))
import k4nerf  # noqa: E402
from oracle import sftnet  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    net = k4nerf.SFTNet(3, 4, 64, 5, 32, 1)
    net.load_state_dict(sftnet.random_state_dict(seed=3, scale=1.0))
    net = net.to(dev)
    res = {}
    for hw in ((520, 520), (398, 520), (256, 256)):
        g = torch.Generator().manual_seed(1)
        x = torch.rand(1, 3, *hw, generator=g).to(dev)
        c = torch.rand(1, 1, *hw, generator=g).to(dev)
        for _ in range(3):
            y = net(x, c)
        torch.cuda.synchronize()

        def timeit(fn, n=10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        eager = timeit(lambda: net(x, c))
        graph = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            net(x, c)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(graph):
            yg = net(x, c)
        graph.replay()
        torch.cuda.synchronize()
        same = bool(torch.equal(y, yg))
        rep = timeit(graph.replay)
        res[f'{hw[0]}x{hw[1]}'] = {'eager_ms': eager, 'graph_ms': rep, 'identical': same}
    print(json.dumps(res))


if __name__ == '__main__':
    main()

"""One decoder tile (default 520x520) twice -- for `ncu --metrics gpu__time_duration.sum` launch lists."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, '4k-nerf_b200')):
    sys.path.insert(0, p)
import k4nerf  # noqa: E402
from oracle import sftnet  # noqa: E402

h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (520, 520)
dev = torch.device('cuda', 0)
net = k4nerf.SFTNet(3, 4, 64, 5, 32, 1)
net.load_state_dict(sftnet.random_state_dict(seed=3, scale=1.0))
net = net.to(dev)
g = torch.Generator().manual_seed(1)
x = torch.rand(1, 3, h, w, generator=g).to(dev)
c = torch.rand(1, 1, h, w, generator=g).to(dev)
for _ in range(2):
    y = net(x, c)
torch.cuda.synchronize()
print('done', float(y.mean()))

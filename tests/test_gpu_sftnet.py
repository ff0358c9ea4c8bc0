"""GPU: the VC-Decoder (SFTNet) conv pipeline through the C ABI against (i) golden vectors produced
by the REFERENCE's own lib/sr_esrnet.py (tests/golden/sftnet_ref.pt) and (ii) the functional
oracle.  Convolution operands are fp16 with fp32 accumulation (the reference's cuDNN convs run
with TF32 operands by default), the trunk / SFT / CondNet math is fp32: bar = 60 dB PSNR."""
import os

import pytest
import torch

import k4nerf
from oracle import pipeline, sftnet

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _inputs():
    g = torch.Generator().manual_seed(11)
    x = torch.rand(1, 3, 20, 28, generator=g) * 1.2 - 0.1
    c = torch.rand(1, 1, 20, 28, generator=g)
    xt = torch.rand(1, 3, 24, 40, generator=g) * 1.2 - 0.1
    ct = torch.rand(1, 24, 40, generator=g)
    return x, c, xt, ct


@pytest.fixture(scope='module')
def net(cuda_device):
    n = k4nerf.SFTNet(3, 4, 64, 5, 32, 1)
    n.load_state_dict(sftnet.random_state_dict(seed=3))
    return n.to(cuda_device)


def test_forward_matches_reference_golden(net, cuda_device):
    gold = torch.load(os.path.join(GOLD, 'sftnet_ref.pt'), map_location='cpu', weights_only=False)
    x, c, _, _ = _inputs()
    y = net(x.to(cuda_device), c.to(cuda_device)).cpu()
    assert y.shape == gold['forward'].shape
    p = pipeline.psnr(y, gold['forward'])
    print('SFTNet forward PSNR vs reference', p, 'maxabs', (y - gold['forward']).abs().max().item(),
          'ref range', gold['forward'].min().item(), gold['forward'].max().item())
    assert p >= 60.0, p


def test_tile_process_matches_reference_golden(net, cuda_device):
    gold = torch.load(os.path.join(GOLD, 'sftnet_ref.pt'), map_location='cpu', weights_only=False)
    _, _, xt, ct = _inputs()
    y = net.tile_process(xt.to(cuda_device), ct.to(cuda_device), tile_size=16, tile_pad=10)
    assert y.device.type == 'cpu' and y.shape == gold['tile_process'].shape
    p = pipeline.psnr(y, gold['tile_process'])
    assert p >= 60.0, p


@pytest.mark.parametrize('hw', [(19, 13), (8, 8), (33, 50), (1, 7)])
def test_ragged_sizes_match_oracle(net, cuda_device, hw):
    """Sizes that are not multiples of the 16x8 pixel tile, down to a single row."""
    h, w = hw
    g = torch.Generator().manual_seed(5)
    x = torch.rand(1, 3, h, w, generator=g)
    c = torch.rand(1, 1, h, w, generator=g)
    sd = sftnet.random_state_dict(seed=3)
    ref = sftnet.sftnet_forward(sd, x, c)
    y = net(x.to(cuda_device), c.to(cuda_device)).cpu()
    p = pipeline.psnr(y, ref)
    assert p >= 60.0, (hw, p)


def test_parameter_update_rebuilds_device_copy(net, cuda_device):
    x, c, _, _ = _inputs()
    a = net(x.to(cuda_device), c.to(cuda_device)).clone()
    with torch.no_grad():
        net.conv_last.bias.add_(0.25)
    b = net(x.to(cuda_device), c.to(cuda_device))
    assert abs((b - a).mean().item() - 0.25) < 1e-3
    with torch.no_grad():
        net.conv_last.bias.sub_(0.25)

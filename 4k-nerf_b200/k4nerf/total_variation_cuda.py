"""Drop-in for the reference's pybind module ``total_variation_cuda`` (lib/cuda/total_variation.cpp:16-24),
backed by csrc/k4_train.cu through the C ABI; bit-identical to the reference extension
(tests/test_gpu_train_ops.py).  Runs on the current stream."""
import ctypes as C

import torch

from . import _lib
from .render_utils_cuda import _chk, _p, _s, _call


def total_variation_add_grad(param, grad, wx, wy, wz, dense_mode):
    """grad += d/dparam of the clamped-difference TV loss of a [1,C,I,J,K] grid (in place)."""
    _chk(param, grad)
    if param.dim() != 5 or grad.shape != param.shape or param.dtype != torch.float32:
        raise RuntimeError('total_variation_add_grad expects float32 [1,C,I,J,K] param and grad of the same shape')
    _call('k4_op_total_variation_add_grad', _p(param), _p(grad), float(wx), float(wy), float(wz), int(bool(dense_mode)),
          param.numel(), int(param.shape[2]), int(param.shape[3]), int(param.shape[4]), _s(param))
    torch.autograd.graph.increment_version(grad)        # written through a raw pointer


__all__ = ['total_variation_add_grad']
del C, _lib

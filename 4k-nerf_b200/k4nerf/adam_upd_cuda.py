"""Drop-in for the reference's pybind module ``adam_upd_cuda`` (lib/cuda/adam_upd.cpp:37-81): the three
in-place Adam updates, backed by csrc/k4_train.cu through the C ABI; bit-identical to the reference
extension (tests/test_gpu_train_ops.py).  Runs on the current stream."""
import torch

from .render_utils_cuda import _chk, _p, _s, _call


def _check(param, grad, exp_avg, exp_avg_sq, *more):
    _chk(param, grad, exp_avg, exp_avg_sq, *more)
    for t in (grad, exp_avg, exp_avg_sq) + more:
        if t.numel() != param.numel() or t.dtype != torch.float32:
            raise RuntimeError('adam update expects float32 tensors of one size')


def _bump(*tensors):
    """The kernels write through raw pointers: tell autograd / the device-scene fingerprint
    (k4nerf._scene._fp uses tensor._version) that the tensors changed in place."""
    for t in tensors:
        torch.autograd.graph.increment_version(t)


def adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    _check(param, grad, exp_avg, exp_avg_sq)
    _call('k4_op_adam_upd', _p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), None, param.numel(), int(step),
          float(beta1), float(beta2), float(lr), float(eps), 0, _s(param))
    _bump(param, exp_avg, exp_avg_sq)


def masked_adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    """Only elements with a non-zero gradient are touched (moments included)."""
    _check(param, grad, exp_avg, exp_avg_sq)
    _call('k4_op_adam_upd', _p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), None, param.numel(), int(step),
          float(beta1), float(beta2), float(lr), float(eps), 1, _s(param))
    _bump(param, exp_avg, exp_avg_sq)


def adam_upd_with_perlr(param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps):
    _check(param, grad, exp_avg, exp_avg_sq, perlr)
    _call('k4_op_adam_upd', _p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), _p(perlr), param.numel(), int(step),
          float(beta1), float(beta2), float(lr), float(eps), 0, _s(param))
    _bump(param, exp_avg, exp_avg_sq)


__all__ = ['adam_upd', 'masked_adam_upd', 'adam_upd_with_perlr']

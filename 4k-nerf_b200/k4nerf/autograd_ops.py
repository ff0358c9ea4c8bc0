"""Autograd shims over the op-level kernels (csrc/k4_ops.cu) with the names the reference's training
code imports from lib/dvgo.py:453-511: ``Raw2Alpha``, ``Raw2Alpha_nonuni``, ``Alphas2Weights``.
Forward and backward values are bit-identical to the reference extension (tests/test_gpu_ops_module.py).
SURVEY.md section 8 f-2."""
import torch
from torch.autograd.function import once_differentiable

from . import render_utils_cuda as _ops


def _raw2alpha_function(name, fwd, bwd, doc):
    """alpha = 1 - (1 + exp(density + shift)) ** (-interval).  d alpha / d density needs
    exp(density + shift), which the forward kernel hands back, so backward never recomputes it."""

    def forward(ctx, density, shift, interval):
        e, alpha = fwd(density, shift, interval)
        if density.requires_grad:
            ctx.interval = interval
            ctx.save_for_backward(e)
        return alpha

    @once_differentiable
    def backward(ctx, grad_alpha):
        (e,) = ctx.saved_tensors
        return bwd(e, grad_alpha.contiguous(), ctx.interval), None, None

    return type(name, (torch.autograd.Function,), {'forward': staticmethod(forward), 'backward': staticmethod(backward),
                                                   '__doc__': doc, '__module__': __name__})


Raw2Alpha = _raw2alpha_function('Raw2Alpha', _ops.raw2alpha, _ops.raw2alpha_backward,
                                'Scalar interval (lib/dvgo.py:453-477).')
Raw2Alpha_nonuni = _raw2alpha_function('Raw2Alpha_nonuni', _ops.raw2alpha_nonuni, _ops.raw2alpha_nonuni_backward,
                                       'Per-sample interval tensor (lib/dvgo.py:479-493).')


class Alphas2Weights(torch.autograd.Function):
    """(weights, alphainv_last) of the flat, ray-sorted alpha list; the serial transmittance scan and its
    suffix-sum backward are the reference's (render_utils_kernel.cu:577-605,654-681)."""

    @staticmethod
    def forward(ctx, alpha, ray_id, N):
        weights, T, last, i_start, i_end = _ops.alpha2weight(alpha, ray_id, N)
        if alpha.requires_grad:
            ctx.n_rays = N
            ctx.save_for_backward(alpha, weights, T, last, i_start, i_end)
        return weights, last

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_weights, grad_last):
        alpha, weights, T, last, i_start, i_end = ctx.saved_tensors
        g = _ops.alpha2weight_backward(alpha, weights, T, last, i_start, i_end, ctx.n_rays,
                                       grad_weights.contiguous(), grad_last.contiguous())
        return g, None, None

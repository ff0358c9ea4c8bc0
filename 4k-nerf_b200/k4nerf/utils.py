"""Checkpoint helpers with the reference names (lib/utils.py:53-66), PSNR (lib/utils.py:18-19) and SSIM (lib/utils.py:88-134)."""
import numpy as np
import torch


def load_model(model_class, ckpt_path):
    """``model_class(**ckpt['model_kwargs']); load_state_dict`` -- the reference's .tar layout
    ``{global_step, model_kwargs, model_state_dict, optimizer_state_dict}`` loads unchanged."""
    ckpt = torch.load(ckpt_path, map_location='cpu', weights_only=False)
    model = model_class(**ckpt['model_kwargs'])
    model.load_state_dict(ckpt['model_state_dict'])
    return model


def mse2psnr(x):
    return -10. * torch.log10(x)


def to8b(x):
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)


def rgb_ssim(img0, img1, max_val, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03, return_map=False):
    """Mean structural similarity of two ``[H, W, 3]`` images as ``lib/utils.py:88-134`` (the mip-NeRF evaluation
    variant) defines it: per channel, local means / variances / covariance under a ``filter_size`` Gaussian window
    (sigma ``filter_sigma``), 'valid' positions only, variances clipped at 0 and the covariance at the Cauchy-Schwarz
    bound, ``c1 = (k1 max_val)^2``, ``c2 = (k2 max_val)^2``.  CPU / numpy: an evaluation metric of
    ``render_viewpoints``, not part of the rendering path."""
    img0 = np.asarray(img0, dtype=np.float64)
    img1 = np.asarray(img1, dtype=np.float64)
    assert img0.ndim == 3 and img0.shape[-1] == 3 and img0.shape == img1.shape
    half = filter_size // 2
    offs = np.arange(filter_size) - half + (2 * half - filter_size + 1) / 2      # centred taps (even sizes: half-pixel shift)
    win = np.exp(-0.5 * (offs / filter_sigma) ** 2)
    win /= win.sum()

    def blur(z):          # separable 'valid' Gaussian window over the two image axes
        z = np.tensordot(np.lib.stride_tricks.sliding_window_view(z, filter_size, axis=0), win[::-1], axes=([-1], [0]))
        return np.tensordot(np.lib.stride_tricks.sliding_window_view(z, filter_size, axis=1), win[::-1], axes=([-1], [0]))

    m0, m1 = blur(img0), blur(img1)
    v0 = np.maximum(blur(img0 * img0) - m0 * m0, 0.0)
    v1 = np.maximum(blur(img1 * img1) - m1 * m1, 0.0)
    cov = blur(img0 * img1) - m0 * m1
    cov = np.sign(cov) * np.minimum(np.sqrt(v0 * v1), np.abs(cov))
    c1, c2 = (k1 * max_val) ** 2, (k2 * max_val) ** 2
    ssim_map = ((2 * m0 * m1 + c1) * (2 * cov + c2)) / ((m0 * m0 + m1 * m1 + c1) * (v0 + v1 + c2))
    return ssim_map if return_map else float(ssim_map.mean())

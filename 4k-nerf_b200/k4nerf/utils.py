"""Checkpoint helpers with the reference names (lib/utils.py:53-66) and PSNR (lib/utils.py:18-19)."""
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

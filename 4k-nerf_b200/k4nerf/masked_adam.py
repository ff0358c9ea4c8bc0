"""MaskedAdam with the reference's interface (lib/masked_adam.py:17-73): Adam with an optional
per-voxel learning rate and a masked update that leaves elements with zero gradient untouched
(moments included), one fused kernel per parameter (csrc/k4_train.cu)."""
import torch

from . import adam_upd_cuda


class MaskedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.99), eps=1e-8):
        if not 0.0 <= lr:
            raise ValueError('Invalid learning rate: {}'.format(lr))
        if not 0.0 <= eps:
            raise ValueError('Invalid epsilon value: {}'.format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError('Invalid beta parameter at index 0: {}'.format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError('Invalid beta parameter at index 1: {}'.format(betas[1]))
        self.per_lr = None
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    def set_pervoxel_lr(self, count):
        assert self.param_groups[0]['params'][0].shape == count.shape
        self.per_lr = (count.float() / count.max()).contiguous()

    @torch.no_grad()
    def step(self):
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            skip_zero_grad = group.get('skip_zero_grad', False)
            for param in group['params']:
                if param.grad is None:
                    continue
                state = self.state[param]
                if len(state) == 0:
                    state['step'] = 0
                    state['exp_avg'] = torch.zeros_like(param, memory_format=torch.preserve_format)
                    state['exp_avg_sq'] = torch.zeros_like(param, memory_format=torch.preserve_format)
                state['step'] += 1
                args = (param, param.grad, state['exp_avg'], state['exp_avg_sq'])
                tail = (state['step'], beta1, beta2, group['lr'], group['eps'])
                if self.per_lr is not None and param.shape == self.per_lr.shape:
                    adam_upd_cuda.adam_upd_with_perlr(*args, self.per_lr, *tail)
                elif skip_zero_grad:
                    adam_upd_cuda.masked_adam_upd(*args, *tail)
                else:
                    adam_upd_cuda.adam_upd(*args, *tail)

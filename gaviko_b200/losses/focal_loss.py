"""Drop-in for reference ``src/losses/focal_loss.py`` (``FocalLoss(gamma)(logits, target)``), as one fused CUDA kernel.

The reference's forward clamps the *logits* to [eps, 1-eps], softmaxes, clamps and softmaxes again
(``focal_loss.py:84-91,94,102``); that exact arithmetic (and its zero gradient outside the clamp range) is what the
kernel reproduces.  The reference also asserts the probabilities lie in [0, 1] with a device->host sync
(``focal_loss.py:95``); a softmax output always does, so no sync is issued here.
"""
from typing import Union

import torch
from torch import Tensor, nn

from .. import ops


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, kind, gamma, eps, ignore_index):
        z = logits.float().contiguous()
        loss, dz = ops.loss_fwd_bwd(z, target.contiguous(), kind, gamma, eps, ignore_index, need_grad=logits.requires_grad)
        ctx.save_for_backward(dz) if dz is not None else None
        ctx.in_dtype = logits.dtype
        return loss.to(logits.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        (dz,) = ctx.saved_tensors
        return (dz * grad_out.float()).to(ctx.in_dtype), None, None, None, None, None


class FocalLoss(nn.Module):
    def __init__(self, gamma, weights: Union[None, Tensor] = None, reduction: str = 'mean', ignore_index=-100, eps=1e-16, fp16: bool = False) -> None:
        super().__init__()
        if reduction not in ['mean', 'none', 'sum']:
            raise NotImplementedError('Reduction {} not implemented.'.format(reduction))
        assert weights is None or isinstance(weights, Tensor), 'weights should be of type Tensor or None, but {} given'.format(type(weights))
        if reduction != 'mean' or weights is not None:
            raise NotImplementedError('gaviko_b200.FocalLoss implements the path the reference trains with: reduction="mean", weights=None')
        self.dtype = torch.float16 if fp16 else torch.float32
        self.reduction = reduction
        self.gamma = gamma
        self.ignore_index = ignore_index
        self.eps = eps
        self.weights = weights

    def forward(self, x: Tensor, target: Tensor) -> Tensor:
        if x.dim() != 2 or x.shape[-1] < 2:
            raise NotImplementedError('gaviko_b200.FocalLoss supports multi-class logits of shape (B, C), C >= 2')
        return _LossFn.apply(x, target.view(-1), ops.LOSS_FOCAL, float(self.gamma), float(self.eps), int(self.ignore_index))


class CrossEntropyLoss(nn.Module):
    """nn.CrossEntropyLoss() alternative of reference ``src/train.py:179`` on the same fused kernel."""

    def forward(self, x: Tensor, target: Tensor) -> Tensor:
        return _LossFn.apply(x, target.view(-1), ops.LOSS_CE, 0.0, 0.0, -100)

"""calc_mse_loss -- drop-in for the reference's src/loss/loss.py:26-46 (the only loss train.py
uses), backed by a single deterministic CUDA reduction that also emits the gradient.
The 16 experimental losses of the reference (TV / Fourier / Huber / ...) are out of scope
(unused by train.py)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import _lib


class _MaskedChunkMse(Function):
    """sum over ray chunks of mean over masked-in rays of (target - pred)^2   (train.py:69-127)."""

    @staticmethod
    def forward(ctx, pred, target, mask, chunk):
        L_ = _lib.lib()
        pred = _lib.require_cuda(pred.contiguous(), "y")
        target = _lib.require_cuda(target.contiguous(), "x")
        if mask is not None:
            mask = _lib.require_cuda(mask.contiguous().to(torch.uint8), "mask", torch.uint8)
        n = pred.numel()
        out = torch.empty(2, device=pred.device, dtype=torch.float32)
        dpred = torch.empty_like(pred)
        with torch.cuda.device(pred.device):
            _lib.check(L_.nafb_mse_loss(_lib.ptr(pred), _lib.ptr(target), _lib.ptr(mask), n, int(chunk or 0), 1.0, _lib.ptr(out),
                                        _lib.ptr(dpred), 0, _lib.stream_ptr()))
        ctx.save_for_backward(dpred)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        gp = g * dpred if ctx.needs_input_grad[0] else None
        gt = -(g * dpred) if ctx.needs_input_grad[1] else None
        return gp, gt, None, None


def masked_chunk_mse(pred, target, mask=None, chunk=None):
    return _MaskedChunkMse.apply(pred.reshape(-1), target.reshape(-1), mask, chunk)


def calc_mse_loss(loss, x, y, tv_loss=None):
    """loss["loss"] += mean((x - y)^2); loss["loss_mse"] = that mean (+ optional tv term)."""
    if x.numel() == 0:
        loss_mse = torch.mean((x - y) ** 2)  # NaN, as in the reference
    else:
        loss_mse = _MaskedChunkMse.apply(y.reshape(-1), x.reshape(-1), None, None)
    loss["loss"] += loss_mse
    loss["loss_mse"] = loss_mse
    if tv_loss is not None:
        loss["loss"] += tv_loss
        loss["tv_loss"] = tv_loss
    return loss


def compute_tv_regularization(loss, values, weight):
    """TV along rays added to the running loss (src/loss/loss.py:10-24)."""
    loss["loss"] = loss["loss"] + torch.sum(torch.abs(values[:, 1:, :] - values[:, :-1, :])) * weight
    return loss

"""Fused photometric loss for the rendered image.

``mse_loss(color, target, weight)`` equals ``weight * ((color - target) ** 2).mean()`` -- the reference's ``LossMse``
(/root/reference/src/loss/loss_mse.py:22-31) -- but computes the value and the seed gradient of the rasterizer's backward
pass in one kernel of libsplatter360.so.  CUDA only."""
from __future__ import annotations

import ctypes

import torch
from torch import Tensor

from . import _lib


class _MseLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, color: Tensor, target: Tensor, weight: float):
        if color.device.type != "cuda":
            raise RuntimeError("splatter360_b200.loss.mse_loss needs CUDA tensors (there is no CPU path)")
        if color.shape != target.shape:
            raise ValueError("color and target must have the same shape")
        lib = _lib.load()
        c = color.contiguous().float()
        t = target.to(c.device).contiguous().float()
        loss = torch.empty((), dtype=torch.float32, device=c.device)
        grad = torch.empty_like(c)
        with torch.cuda.device(c.device):
            _lib.check(lib.s360_mse_loss_grad(
                ctypes.c_void_p(c.data_ptr()), ctypes.c_void_p(t.data_ptr()), ctypes.c_int64(c.numel()),
                ctypes.c_float(weight), ctypes.c_void_p(loss.data_ptr()), ctypes.c_void_p(grad.data_ptr()),
                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        return grad * grad_out, None, None


def mse_loss(color: Tensor, target: Tensor, weight: float = 1.0) -> Tensor:
    return _MseLoss.apply(color, target, float(weight))

"""Fused BatchNorm(+ReLU) for channels-last bf16 activations (``csrc/bn.cu``).

``FusedBatchNorm2d`` is a drop-in ``nn.BatchNorm2d`` subclass (same parameters, buffers and
``state_dict`` keys, so reference / torchvision checkpoints load unchanged).  On CUDA, when the
input is a channels-last bf16 tensor (what a conv produces under bf16 autocast), forward and
backward run the hand-written streaming kernels with the following ReLU folded in; every other
case (CPU, fp32, NCHW) takes the stock ``F.batch_norm`` path, which is also the test oracle.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _stream, count_launch, require_ext, sm_count


def _eligible(x: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[1] % 8 == 0
            and x.shape[1] <= 2048 and x.is_contiguous(memory_format=torch.channels_last)
            and x.numel() > 0)


def _direct_grad_ptrs(mod, weight, bias):
    """Arena pointers for dgamma / dbeta when the layer is in direct-gradient mode: the backward
    kernel then writes the parameter gradients in place (no AccumulateGrad add, no temporaries)."""
    if mod is None or not getattr(mod, "_direct_grad", False) or weight is None or bias is None:
        return None
    gw, gb = weight.grad, bias.grad
    if (gw is None or gb is None or gw.dtype != torch.float32 or gb.dtype != torch.float32
            or not gw.is_contiguous() or not gb.is_contiguous() or gw.device != weight.device):
        return None
    return gw.data_ptr(), gb.data_ptr()


class _FusedBN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual, running_mean, running_var, training, momentum, eps, relu,
                mod, nbt):
        ext = require_ext()
        N, C, H, W = x.shape
        R = N * H * W
        dev = x.device
        sms = sm_count(dev)
        y = torch.empty_like(x, memory_format=torch.channels_last)
        stats = torch.empty((4, C), dtype=torch.float32, device=dev)  # mean, invstd, scale, shift
        mean, invstd, scale, shift = stats[0], stats[1], stats[2], stats[3]
        if training:
            nb = ext.bn_partial_blocks(R, sms)
            partial = torch.empty((nb, C, 2), dtype=torch.float32, device=dev)
            pptr = partial.data_ptr()
        else:
            invstd.copy_(torch.rsqrt(running_var + eps))
            mean.copy_(running_mean)
            g = weight if weight is not None else torch.ones_like(invstd)
            b = bias if bias is not None else torch.zeros_like(invstd)
            scale.copy_(g * invstd)
            shift.copy_(b - running_mean * scale)
            pptr = 0
        count_launch(ext.bn_forward(x.data_ptr(), y.data_ptr(), R, C,
                       weight.data_ptr() if weight is not None else 0,
                       bias.data_ptr() if bias is not None else 0,
                       running_mean.data_ptr() if (training and running_mean is not None) else 0,
                       running_var.data_ptr() if (training and running_var is not None) else 0,
                       mean.data_ptr(), invstd.data_ptr(), scale.data_ptr(), shift.data_ptr(), pptr,
                       float(eps), float(momentum), int(relu), int(training),
                       residual.data_ptr() if residual is not None else 0,
                       nbt.data_ptr() if (training and nbt is not None) else 0, sms, _stream(dev)))
        ctx.has_res = residual is not None
        # with a residual the ReLU mask cannot be recomputed from x alone: keep the output (the next
        # layer saves it anyway, so this costs no memory)
        ctx.save_for_backward(x, weight, stats, y if (ctx.has_res and relu) else None)
        ctx.relu = bool(relu)
        ctx.training = bool(training)
        ctx.mod = mod
        ctx.bias = bias
        return y

    @staticmethod
    def backward(ctx, dy):
        ext = require_ext()
        x, weight, stats, ysaved = ctx.saved_tensors
        N, C, H, W = x.shape
        R = N * H * W
        dev = x.device
        sms = sm_count(dev)
        dy = dy.contiguous(memory_format=torch.channels_last)
        if dy.dtype != torch.bfloat16:
            dy = dy.to(torch.bfloat16)
        mean, invstd, scale, shift = stats[0], stats[1], stats[2], stats[3]
        nones = (None,) * 8
        if not ctx.training:
            # eval mode: statistics are constants -> dx = dy' * scale
            d = dy.float()
            if ctx.relu:
                if ysaved is not None:
                    d = d * (ysaved > 0)
                else:
                    d = d * ((x.float() * scale.view(1, C, 1, 1) + shift.view(1, C, 1, 1)) > 0)
            dxe = (d * scale.view(1, C, 1, 1)).to(torch.bfloat16)
            xhat = (x.float() - mean.view(1, C, 1, 1)) * invstd.view(1, C, 1, 1)
            dg = (d * xhat).sum(dim=(0, 2, 3)) if weight is not None else None
            db = d.sum(dim=(0, 2, 3)) if weight is not None else None
            dres = d.to(torch.bfloat16) if ctx.has_res else None
            return (dxe, dg, db, dres) + nones
        dx = torch.empty_like(x, memory_format=torch.channels_last)
        nb = ext.bn_partial_blocks(R, sms)
        partial = torch.empty((nb, C, 2), dtype=torch.float32, device=dev)
        direct = _direct_grad_ptrs(ctx.mod, weight, ctx.bias)
        grads = torch.empty((3 if direct else 5, C), dtype=torch.float32, device=dev)  # coef[3] (+ dgamma, dbeta)
        if direct:
            dg_ptr, db_ptr = direct
        else:
            dg_ptr, db_ptr = grads[3].data_ptr(), grads[4].data_ptr()
        dres = None
        dres_ptr = 0
        if ctx.has_res:
            if ctx.relu:
                dres = torch.empty_like(x, memory_format=torch.channels_last)
                dres_ptr = dres.data_ptr()
            else:
                dres = dy
        count_launch(ext.bn_backward(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), R, C,
                        weight.data_ptr() if weight is not None else 0, mean.data_ptr(), invstd.data_ptr(),
                        scale.data_ptr(), shift.data_ptr(), partial.data_ptr(), dg_ptr, db_ptr,
                        grads[0].data_ptr(), int(ctx.relu), ysaved.data_ptr() if ysaved is not None else 0,
                        dres_ptr, sms, _stream(dev)))
        if direct or weight is None:
            return (dx, None, None, dres) + nones
        return (dx, grads[3], grads[4], dres) + nones


class FusedBatchNorm2d(nn.BatchNorm2d):
    """``nn.BatchNorm2d`` with an optional fused residual add + ReLU and hand-written sm_100a
    kernels: ``forward(x, residual=None)`` returns ``[relu](bn(x) [+ residual])``."""

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float = 0.1, affine: bool = True,
                 track_running_stats: bool = True, relu: bool = False, device=None, dtype=None):
        super().__init__(num_features, eps=eps, momentum=momentum, affine=affine,
                         track_running_stats=track_running_stats, device=device, dtype=dtype)
        self.fused_relu = relu

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        use_batch_stats = self.training or not self.track_running_stats
        if (_eligible(x) and self.momentum is not None
                and (use_batch_stats or self.running_mean is not None)
                and (residual is None or (_eligible(residual) and residual.shape == x.shape))):
            # num_batches_tracked is bumped inside the statistics kernel (no extra launch)
            nbt = (self.num_batches_tracked
                   if (self.training and self.track_running_stats and self.num_batches_tracked is not None
                       and self.num_batches_tracked.dtype == torch.int64) else None)
            return _FusedBN.apply(x, self.weight, self.bias, residual,
                                  self.running_mean if self.track_running_stats else None,
                                  self.running_var if self.track_running_stats else None,
                                  use_batch_stats, self.momentum, self.eps, self.fused_relu, self, nbt)
        y = super().forward(x)
        if residual is not None:
            y = y + residual
        return F.relu(y) if self.fused_relu else y


__all__ = ["FusedBatchNorm2d"]

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

from . import _stream, require_ext, sm_count


def _eligible(x: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[1] % 8 == 0
            and x.shape[1] <= 2048 and x.is_contiguous(memory_format=torch.channels_last)
            and x.numel() > 0)


class _FusedBN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, training, momentum, eps, relu):
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
        ext.bn_forward(x.data_ptr(), y.data_ptr(), R, C,
                       weight.data_ptr() if weight is not None else 0,
                       bias.data_ptr() if bias is not None else 0,
                       running_mean.data_ptr() if (training and running_mean is not None) else 0,
                       running_var.data_ptr() if (training and running_var is not None) else 0,
                       mean.data_ptr(), invstd.data_ptr(), scale.data_ptr(), shift.data_ptr(), pptr,
                       float(eps), float(momentum), int(relu), int(training), sms, _stream(dev))
        ctx.save_for_backward(x, weight, stats)
        ctx.relu = bool(relu)
        ctx.training = bool(training)
        return y

    @staticmethod
    def backward(ctx, dy):
        ext = require_ext()
        x, weight, stats = ctx.saved_tensors
        N, C, H, W = x.shape
        R = N * H * W
        dev = x.device
        sms = sm_count(dev)
        dy = dy.contiguous(memory_format=torch.channels_last)
        if dy.dtype != torch.bfloat16:
            dy = dy.to(torch.bfloat16)
        dx = torch.empty_like(x, memory_format=torch.channels_last)
        mean, invstd, scale, shift = stats[0], stats[1], stats[2], stats[3]
        if not ctx.training:
            # eval mode: statistics are constants -> dx = dy' * scale
            z = x.float() * scale.view(1, C, 1, 1) + shift.view(1, C, 1, 1)
            d = dy.float() * ((z > 0) if ctx.relu else 1.0)
            dxe = (d * scale.view(1, C, 1, 1)).to(torch.bfloat16)
            xhat = (x.float() - mean.view(1, C, 1, 1)) * invstd.view(1, C, 1, 1)
            dg = (d * xhat).sum(dim=(0, 2, 3)) if weight is not None else None
            db = d.sum(dim=(0, 2, 3)) if weight is not None else None
            return dxe, dg, db, None, None, None, None, None, None
        nb = ext.bn_partial_blocks(R, sms)
        partial = torch.empty((nb, C, 2), dtype=torch.float32, device=dev)
        grads = torch.empty((5, C), dtype=torch.float32, device=dev)  # dgamma, dbeta, coef[3]
        ext.bn_backward(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), R, C,
                        weight.data_ptr() if weight is not None else 0, mean.data_ptr(), invstd.data_ptr(),
                        scale.data_ptr(), shift.data_ptr(), partial.data_ptr(), grads[0].data_ptr(),
                        grads[1].data_ptr(), grads[2].data_ptr(), int(ctx.relu), sms, _stream(dev))
        dg = grads[0] if weight is not None else None
        db = grads[1] if weight is not None else None
        return dx, dg, db, None, None, None, None, None, None


class FusedBatchNorm2d(nn.BatchNorm2d):
    """``nn.BatchNorm2d`` with an optional fused ReLU and hand-written sm_100a kernels."""

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float = 0.1, affine: bool = True,
                 track_running_stats: bool = True, relu: bool = False, device=None, dtype=None):
        super().__init__(num_features, eps=eps, momentum=momentum, affine=affine,
                         track_running_stats=track_running_stats, device=device, dtype=dtype)
        self.fused_relu = relu

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        use_batch_stats = self.training or not self.track_running_stats
        if (_eligible(x) and self.momentum is not None
                and (use_batch_stats or self.running_mean is not None)):
            if self.training and self.track_running_stats and self.num_batches_tracked is not None:
                self.num_batches_tracked.add_(1)
            return _FusedBN.apply(x, self.weight, self.bias,
                                  self.running_mean if self.track_running_stats else None,
                                  self.running_var if self.track_running_stats else None,
                                  use_batch_stats, self.momentum, self.eps, self.fused_relu)
        y = super().forward(x)
        return F.relu(y) if self.fused_relu else y


__all__ = ["FusedBatchNorm2d"]

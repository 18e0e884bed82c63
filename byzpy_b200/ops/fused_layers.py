"""Replica layers that produce their gradients *in place* in the flat gradient arena.

In the fused parameter-server round a worker's gradient row lives in symmetric memory and every
``param.grad`` is a view into it (``parallel/arena.py``).  Stock autograd then spends a large part
of a batch-32 ResNet step on bookkeeping kernels: a bf16 cast + layout change of every conv weight,
a bf16->fp32 cast of every weight gradient, and an ``add_`` per parameter to accumulate into the
arena view (profiles/bench_log.md: ~175 of ~450 launches per replica step).  The layers here keep
torch.nn's module API and ``state_dict`` layout but, once :func:`enable_direct_grads` has switched
them on, they

* refresh a persistent channels-last bf16 shadow of the weight with one cast kernel,
* call the cuDNN/cuBLAS forward/backward primitives directly,
* write the weight gradient straight into the arena view (one cast-copy) and return ``None`` to
  autograd, and
* run the weight-gradient GEMMs on a *side stream*: the data-gradient chain is the critical path
  of backward, the weight gradients are only needed when the aggregation kernel starts, so they
  overlap with the rest of the backward pass (fork/join by events, CUDA-graph capturable).

Off (the default) they behave exactly like ``nn.Conv2d`` / ``nn.Linear`` / ``nn.MaxPool2d``.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _stream, count_launch, require_ext, sm_count

_aten = torch.ops.aten


class GradSink:
    """Per-replica bookkeeping for direct gradients: the side stream the weight-gradient work runs
    on and the tensors that must stay alive until it has been joined."""

    def __init__(self, side_stream: Optional["torch.cuda.Stream"] = None):
        self.side_stream = side_stream
        self.keep: List[object] = []
        self._forked = False
        self.convs: List["ArenaConv2d"] = []
        self.shadows_fresh = False

    def refresh_shadows(self) -> None:
        """Refresh the bf16 channels-last shadow of EVERY conv weight with one launch
        (``csrc/layout.cu``); the forward passes of this step then skip their own refresh."""
        convs = [m for m in self.convs if m._direct_grad and m.weight.is_cuda]
        if not convs:
            return
        ext = require_ext()
        dev = convs[0].weight.device
        n = ext.krsc_cast([m.weight.data_ptr() for m in convs], [m._shadow().data_ptr() for m in convs],
                          [m.weight.shape[0] for m in convs], [m.weight.shape[1] for m in convs],
                          [m.weight.shape[2] * m.weight.shape[3] for m in convs], 0, _stream(dev))
        count_launch(n)
        self.shadows_fresh = True

    def fork(self) -> Optional["torch.cuda.Stream"]:
        """Make the side stream wait for everything enqueued on the current stream so far."""
        side = self.side_stream
        if side is None:
            return None
        side.wait_stream(torch.cuda.current_stream())
        self._forked = True
        return side

    def join(self) -> None:
        """Order the current stream after all side-stream work of this backward pass."""
        if self._forked:
            torch.cuda.current_stream().wait_stream(self.side_stream)
            self._forked = False
        self.keep.clear()
        self.shadows_fresh = False


def _bf16_autocast_on() -> bool:
    return torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16


def _grad_view_ok(p: Optional[torch.Tensor]) -> bool:
    if p is None:
        return True
    g = p.grad
    return (g is not None and g.dtype == torch.float32 and g.is_contiguous() and g.device == p.device
            and g.shape == p.shape)


# --------------------------------------------------------------------------- convolution
def _krsc_cast(src: torch.Tensor, dst: torch.Tensor, *, to_grad: bool) -> None:
    """fp32 OIHW <-> bf16 channels-last for one filter bank, coalesced on both sides."""
    K, C, R, S = dst.shape
    cl = src if to_grad else dst
    if not cl.is_contiguous(memory_format=torch.channels_last) or (dst if to_grad else src).stride() != (
            C * R * S, R * S, S, 1):
        dst.copy_(src)          # unusual strides (e.g. a 1-channel filter): ATen's strided copy
        return
    ext = require_ext()
    count_launch(ext.krsc_cast([src.data_ptr()], [dst.data_ptr()], [K], [C], [R * S], int(to_grad),
                               _stream(dst.device)))


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, mod):
        w16 = mod._shadow()
        sink = mod._sink
        if sink is None or not sink.shadows_fresh:
            _krsc_cast(weight, w16, to_grad=False)   # fp32 OIHW -> bf16 channels-last
        with torch.autocast("cuda", enabled=False):
            y = _aten.convolution(x, w16, None, mod.stride, mod.padding, mod.dilation, False, (0, 0),
                                  mod.groups)
        ctx.save_for_backward(x)
        ctx.w16 = w16                          # plain attribute: the shadow is overwritten next step
        ctx.mod = mod
        ctx.weight = weight
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        mod, w16, weight = ctx.mod, ctx.w16, ctx.weight
        dy = dy.contiguous(memory_format=torch.channels_last)
        need_dx = ctx.needs_input_grad[0]
        sink: Optional[GradSink] = getattr(mod, "_sink", None)
        args = (None, mod.stride, mod.padding, mod.dilation, False, (0, 0), mod.groups)
        side = sink.fork() if sink is not None else None
        if side is not None:
            with torch.cuda.stream(side):
                gw = _aten.convolution_backward(dy, x, w16, *args, [False, True, False])[1]
                _krsc_cast(gw, weight.grad, to_grad=True)   # cast + layout change into the arena row
            sink.keep.append((dy, x, gw))      # alive until GradSink.join()
            dx = (_aten.convolution_backward(dy, x, w16, *args, [True, False, False])[0]
                  if need_dx else None)
        else:
            dx, gw, _ = _aten.convolution_backward(dy, x, w16, *args, [need_dx, True, False])
            _krsc_cast(gw, weight.grad, to_grad=True)
        return dx, None, None


class ArenaConv2d(nn.Conv2d):
    """``nn.Conv2d`` whose weight gradient can be produced in place (see module docstring)."""

    _direct_grad = False
    _sink: Optional[GradSink] = None

    def _shadow(self) -> torch.Tensor:
        w = self.weight
        s = getattr(self, "_w16", None)
        if s is None or s.device != w.device or s.shape != w.shape:
            s = torch.empty(w.shape, dtype=torch.bfloat16, device=w.device, memory_format=torch.channels_last)
            self._w16 = s
        return s

    def _fast(self, x: torch.Tensor) -> bool:
        return (self._direct_grad and x.is_cuda and x.dim() == 4 and self.bias is None
                and self.padding_mode == "zeros" and not isinstance(self.padding, str)
                and torch.is_grad_enabled() and _bf16_autocast_on() and _grad_view_ok(self.weight))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self._fast(x):
            return super().forward(x)
        if x.dtype != torch.bfloat16:
            x = x.to(torch.bfloat16)
        x = x.contiguous(memory_format=torch.channels_last)
        return _ConvFn.apply(x, self.weight, self)


# --------------------------------------------------------------------------- linear
class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, mod):
        w16 = mod._shadow()
        w16.copy_(weight)
        with torch.autocast("cuda", enabled=False):
            y = torch.addmm(bias.to(torch.bfloat16), x, w16.t()) if bias is not None else x @ w16.t()
        ctx.save_for_backward(x)
        ctx.w16, ctx.mod, ctx.weight, ctx.bias = w16, mod, weight, bias
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        mod, w16, weight, bias = ctx.mod, ctx.w16, ctx.weight, ctx.bias
        dy = dy.contiguous()
        sink: Optional[GradSink] = getattr(mod, "_sink", None)
        side = sink.fork() if sink is not None else None

        def param_grads():
            gw = dy.t() @ x
            weight.grad.copy_(gw)
            if bias is not None:
                bias.grad.copy_(dy.sum(0, dtype=torch.float32))
            return gw

        if side is not None:
            with torch.cuda.stream(side):
                gw = param_grads()
            sink.keep.append((dy, x, gw))
        else:
            param_grads()
        dx = dy @ w16 if ctx.needs_input_grad[0] else None
        return dx, None, None, None


class ArenaLinear(nn.Linear):
    """``nn.Linear`` with in-place parameter gradients (2-D inputs; see module docstring)."""

    _direct_grad = False
    _sink: Optional[GradSink] = None

    def _shadow(self) -> torch.Tensor:
        w = self.weight
        s = getattr(self, "_w16", None)
        if s is None or s.device != w.device or s.shape != w.shape:
            s = torch.empty(w.shape, dtype=torch.bfloat16, device=w.device)
            self._w16 = s
        return s

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not (self._direct_grad and x.is_cuda and x.dim() == 2 and torch.is_grad_enabled()
                and _bf16_autocast_on() and _grad_view_ok(self.weight) and _grad_view_ok(self.bias)):
            return super().forward(x)
        if x.dtype != torch.bfloat16:
            x = x.to(torch.bfloat16)
        return _LinearFn.apply(x.contiguous(), self.weight, self.bias, self)


# --------------------------------------------------------------------------- max pooling
class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ext = require_ext()
        N, C, H, W = x.shape
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        dev = x.device
        y = torch.empty((N, C, Ho, Wo), dtype=x.dtype, device=dev, memory_format=torch.channels_last)
        idx = torch.empty((N, Ho, Wo, C), dtype=torch.uint8, device=dev)
        ext.maxpool_forward(x.data_ptr(), y.data_ptr(), idx.data_ptr(), N, H, W, C, sm_count(dev), _stream(dev))
        count_launch()
        ctx.save_for_backward(idx)
        ctx.shape = (N, C, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        ext = require_ext()
        (idx,) = ctx.saved_tensors
        N, C, H, W = ctx.shape
        dev = dy.device
        dy = dy.contiguous(memory_format=torch.channels_last)
        if dy.dtype != torch.bfloat16:
            dy = dy.to(torch.bfloat16)
        dx = torch.empty((N, C, H, W), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last)
        ext.maxpool_backward(dy.data_ptr(), idx.data_ptr(), dx.data_ptr(), N, H, W, C, sm_count(dev),
                             _stream(dev))
        count_launch()
        return dx


class FusedMaxPool2d(nn.MaxPool2d):
    """``nn.MaxPool2d``; the 3x3 / stride 2 / pad 1 case on channels-last bf16 CUDA activations runs
    the hand-written streaming kernels (``csrc/pool.cu``), everything else takes ATen's path."""

    def _fast(self, x: torch.Tensor) -> bool:
        def two(v):
            return tuple(v) if isinstance(v, (tuple, list)) else (v, v)

        return (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[1] % 8 == 0
                and x.numel() > 0 and x.is_contiguous(memory_format=torch.channels_last)
                and two(self.kernel_size) == (3, 3) and two(self.stride) == (2, 2)
                and two(self.padding) == (1, 1) and two(self.dilation) == (1, 1)
                and not self.ceil_mode and not self.return_indices)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self._fast(x):
            return _MaxPoolFn.apply(x)
        return super().forward(x)


# --------------------------------------------------------------------------- switch
def enable_direct_grads(module: nn.Module, *, side_stream: Optional["torch.cuda.Stream"] = None,
                        enabled: bool = True) -> GradSink:
    """Switch every in-place-gradient layer of ``module`` on (or off).

    Call after the parameters' ``.grad`` have been bound to the gradient arena.  In this mode a
    backward pass OVERWRITES the gradients of those layers (the arena is a per-step buffer), and the
    caller must invoke ``sink.join()`` after ``loss.backward()`` before reading the gradients.
    """
    from .fused_bn import FusedBatchNorm2d

    sink = GradSink(side_stream if enabled else None)
    for m in module.modules():
        if isinstance(m, (ArenaConv2d, ArenaLinear)):
            m._direct_grad = enabled
            m._sink = sink if enabled else None
            if isinstance(m, ArenaConv2d) and enabled:
                sink.convs.append(m)
        elif isinstance(m, FusedBatchNorm2d):
            m._direct_grad = enabled
    return sink


__all__ = ["ArenaConv2d", "ArenaLinear", "FusedMaxPool2d", "GradSink", "enable_direct_grads"]

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

import contextlib
from typing import List, Optional

import torch
import torch.nn as nn

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
        self.pending: List[tuple] = []      # (bf16 channels-last wgrad, fp32 arena view) awaiting the cast
        self.shadow_event: Optional["torch.cuda.Event"] = None   # refresh issued on the side stream
        self.flush_every = 6                # wgrads cast per launch while backward is still running

    def refresh_shadows(self) -> None:
        """Refresh the bf16 channels-last shadow of EVERY conv weight with one launch
        (``csrc/layout.cu``); the forward passes of this step then skip their own refresh."""
        convs = [m for m in self.convs if m._direct_grad and m.weight.is_cuda
                 and not (isinstance(m, S2DStemConv2d) and m._s2d_ok())]   # the stem packs its own weight
        if not convs:
            return
        ext = require_ext()
        dev = convs[0].weight.device
        shadows = [m._shadow() for m in convs]          # (allocated on the launching stream, once)
        side = self.side_stream
        # With a side stream the cast overlaps the input pipeline and the stem (which packs its own
        # weight): the first convolution that needs a shadow waits for `shadow_event`.
        ctx = torch.cuda.stream(side) if side is not None else contextlib.nullcontext()
        if side is not None:
            side.wait_stream(torch.cuda.current_stream(dev))     # weights were updated on this stream
        with ctx:
            n = ext.krsc_cast([m.weight.data_ptr() for m in convs], [t.data_ptr() for t in shadows],
                              [m.weight.shape[0] for m in convs], [m.weight.shape[1] for m in convs],
                              [m.weight.shape[2] * m.weight.shape[3] for m in convs], 0, _stream(dev))
            if side is not None:
                self.shadow_event = torch.cuda.Event()
                self.shadow_event.record(side)
        count_launch(n)
        for m in convs:
            m._shadow_fresh = True
        self.shadows_fresh = True

    def wait_shadows(self) -> None:
        """Order the current stream after the asynchronous shadow refresh (first consumer only)."""
        if self.shadow_event is not None:
            torch.cuda.current_stream().wait_event(self.shadow_event)
            self.shadow_event = None

    def flush_pending(self, force: bool = False) -> None:
        """Cast the weight gradients produced so far into the arena row (on the stream that produced
        them).  Called every few convolutions during backward so that only a short tail is left
        after the last weight-gradient GEMM."""
        if not self.pending or (not force and len(self.pending) < self.flush_every):
            return
        ctx = torch.cuda.stream(self.side_stream) if self._forked else contextlib.nullcontext()
        with ctx:
            _krsc_cast_many([g for g, _ in self.pending], [d for _, d in self.pending], to_grad=True)
        self.pending.clear()

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
        self.flush_pending(force=True)      # the (short) tail of weight gradients still to be cast
        self.shadow_event = None
        if self._forked:
            torch.cuda.current_stream().wait_stream(self.side_stream)
            self._forked = False
        self.keep.clear()
        if self.shadows_fresh:
            for m in self.convs:
                m._shadow_fresh = False
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


def _krsc_cast_many(srcs, dsts, *, to_grad: bool) -> None:
    ok = []
    for src, dst in zip(srcs, dsts):
        K, C, R, S = dst.shape
        cl = src if to_grad else dst
        if cl.is_contiguous(memory_format=torch.channels_last) and (dst if to_grad else src).stride() == (
                C * R * S, R * S, S, 1):
            ok.append((src, dst))
        else:
            dst.copy_(src)
    if ok:
        ext = require_ext()
        count_launch(ext.krsc_cast([s.data_ptr() for s, _ in ok], [d.data_ptr() for _, d in ok],
                                   [d.shape[0] for _, d in ok], [d.shape[1] for _, d in ok],
                                   [d.shape[2] * d.shape[3] for _, d in ok], int(to_grad),
                                   _stream(ok[0][1].device)))


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, mod):
        w16 = mod._shadow()
        if not getattr(mod, "_shadow_fresh", False):
            _krsc_cast(weight, w16, to_grad=False)   # fp32 OIHW -> bf16 channels-last
        elif mod._sink is not None:
            mod._sink.wait_shadows()                 # refresh runs asynchronously on the side stream
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
            sink.pending.append((gw, weight.grad))   # cast into the arena row in batches / by join()
            sink.keep.append((dy, x, gw))      # alive until GradSink.join()
            sink.flush_pending()
            dx = (_aten.convolution_backward(dy, x, w16, *args, [True, False, False])[0]
                  if need_dx else None)
        else:
            dx, gw, _ = _aten.convolution_backward(dy, x, w16, *args, [need_dx, True, False])
            if sink is not None:
                sink.pending.append((gw, weight.grad))
                sink.flush_pending()
            else:
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


# --------------------------------------------------------------------------- space-to-depth stem
class PackedStemInput:
    """A 2x2 space-to-depth packed, zero-padded image batch for :class:`S2DStemConv2d`:
    ``data`` is bf16 ``[N, 16, (H-1)//2 + 4, (W-1)//2 + 4]`` channels-last (``csrc/layout.cu``)."""

    __slots__ = ("data", "hw")

    def __init__(self, data: torch.Tensor, hw):
        self.data, self.hw = data, tuple(hw)

    @property
    def shape(self):
        return torch.Size((self.data.shape[0], 3) + self.hw)

    @property
    def device(self):
        return self.data.device


def pack_stem_input(x: torch.Tensor, mean=None, std=None) -> PackedStemInput:
    """uint8 ``[N, H, W, 3]`` (normalised with ``mean`` / ``std``) or bf16 channels-last ``[N, 3, H, W]``
    -> :class:`PackedStemInput`, one streaming kernel."""
    ext = require_ext()
    if x.dtype == torch.uint8:
        N, H, W, C = x.shape
        src, is_u8 = x, 1
        m = [float(v) for v in (mean if isinstance(mean, (list, tuple)) else [127.5 if mean is None else mean] * 3)]
        sd = [float(v) for v in (std if isinstance(std, (list, tuple)) else [127.5 if std is None else std] * 3)]
        sc = [1.0 / v for v in sd]
    else:
        N, C, H, W = x.shape
        src, is_u8 = x, 0
        m, sc = [0.0] * 3, [1.0] * 3
        if x.dtype != torch.bfloat16 or not x.is_contiguous(memory_format=torch.channels_last):
            raise TypeError("pack_stem_input expects uint8 NHWC or bf16 channels-last input")
    if C != 3 or not src.is_contiguous(memory_format=torch.contiguous_format if is_u8 else torch.channels_last):
        raise ValueError("pack_stem_input needs a dense 3-channel image batch")
    Hb, Wb = (H - 1) // 2 + 4, (W - 1) // 2 + 4
    out = torch.empty((N, 16, Hb, Wb), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    ext.s2d_pack(src.data_ptr(), is_u8, out.data_ptr(), N, H, W, m, sc, _stream(x.device))
    count_launch()
    return PackedStemInput(out, (H, W))


class _StemFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xp, weight, mod):
        ext = require_ext()
        K = weight.shape[0]
        wp = mod._packed_weight()
        ext.stem_weight_pack(weight.data_ptr(), wp.data_ptr(), K, _stream(weight.device))
        count_launch()
        with torch.autocast("cuda", enabled=False):
            y = _aten.convolution(xp, wp, None, (1, 1), (0, 0), (1, 1), False, (0, 0), 1)
        ctx.save_for_backward(xp)
        ctx.wp, ctx.mod, ctx.weight = wp, mod, weight
        return y

    @staticmethod
    def backward(ctx, dy):
        ext = require_ext()
        (xp,) = ctx.saved_tensors
        mod, wp, weight = ctx.mod, ctx.wp, ctx.weight
        dy = dy.contiguous(memory_format=torch.channels_last)
        if dy.dtype != torch.bfloat16:
            dy = dy.to(torch.bfloat16)
        direct = mod._direct_grad and _grad_view_ok(weight)
        sink: Optional[GradSink] = getattr(mod, "_sink", None) if direct else None
        target = weight.grad if direct else torch.empty_like(weight, memory_format=torch.contiguous_format)

        def wgrad():
            gp = _aten.convolution_backward(dy, xp, wp, None, (1, 1), (0, 0), (1, 1), False, (0, 0), 1,
                                            [False, True, False])[1]
            gp = gp.contiguous(memory_format=torch.channels_last)
            ext.stem_grad_unpack(gp.data_ptr(), target.data_ptr(), weight.shape[0], _stream(weight.device))
            count_launch()
            return gp

        side = sink.fork() if sink is not None else None
        if side is not None:
            with torch.cuda.stream(side):
                gp = wgrad()
            sink.keep.append((dy, xp, gp))
        else:
            wgrad()
        # the image batch needs no gradient; a packed input never requires grad
        return None, (None if direct else target), None


class S2DStemConv2d(ArenaConv2d):
    """The ResNet stem (7x7 / stride 2 / pad 3 over 3 channels) executed as a 4x4 / stride 1
    convolution over the 2x2 space-to-depth packed input: same parameters, ``state_dict`` and
    results, 4-5x faster on B200 because the dense 16-channel form runs on the sm_100 tensor-core
    kernels (``csrc/layout.cu`` has the index algebra).  Accepts a :class:`PackedStemInput`
    (from ``ops.normalize_uint8_nhwc(..., s2d=True)``) or an ordinary image batch."""

    def _s2d_ok(self) -> bool:
        return (self.kernel_size == (7, 7) and self.stride == (2, 2) and self.padding == (3, 3)
                and self.dilation == (1, 1) and self.groups == 1 and self.in_channels == 3
                and self.bias is None and self.padding_mode == "zeros")

    def _packed_weight(self) -> torch.Tensor:
        w = self.weight
        s = getattr(self, "_wp", None)
        if s is None or s.device != w.device or s.shape[0] != w.shape[0]:
            s = torch.empty((w.shape[0], 16, 4, 4), dtype=torch.bfloat16, device=w.device,
                            memory_format=torch.channels_last)
            self._wp = s
        return s

    def forward(self, x):
        if isinstance(x, PackedStemInput):
            if not self._s2d_ok():
                raise TypeError("PackedStemInput needs a 7x7/s2/p3 3-channel stem convolution")
            return _StemFn.apply(x.data, self.weight, self)
        if (self._direct_grad and self._s2d_ok() and x.is_cuda and x.dim() == 4 and x.shape[1] == 3
                and torch.is_grad_enabled() and not x.requires_grad and _bf16_autocast_on()):
            xb = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)
            xb = xb.contiguous(memory_format=torch.channels_last)
            return _StemFn.apply(pack_stem_input(xb).data, self.weight, self)
        return super().forward(x)


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
    """``nn.Linear`` with in-place parameter gradients (leading dimensions are flattened; see module
    docstring)."""

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
        if not (self._direct_grad and x.is_cuda and x.dim() >= 2 and torch.is_grad_enabled()
                and _bf16_autocast_on() and _grad_view_ok(self.weight) and _grad_view_ok(self.bias)):
            return super().forward(x)
        if x.dtype != torch.bfloat16:
            x = x.to(torch.bfloat16)
        lead = x.shape[:-1]
        y = _LinearFn.apply(x.reshape(-1, x.shape[-1]).contiguous(), self.weight, self.bias, self)
        return y.view(*lead, y.shape[-1])


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


# --------------------------------------------------------------------------- gradient buckets
class _BucketMark(torch.autograd.Function):
    """Identity in forward.  Its backward runs when the gradient with respect to this activation
    exists, i.e. when every layer downstream of it has finished its backward: all parameter
    gradients of those layers have been issued (AccumulateGrad nodes run before lower-priority
    nodes, direct-gradient layers enqueue theirs inside their own backward).  The callback turns
    that moment into CUDA events the round engine's aggregation stream waits on."""

    @staticmethod
    def forward(ctx, x, callback, index):
        ctx.callback, ctx.index = callback, index
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        ctx.callback(ctx.index)
        return g, None, None


def install_bucket_marks(marks, callback):
    """``marks`` = [(module, bucket index)]: wrap each module's first positional input in a
    :class:`_BucketMark`.  Returns the hook handles (``h.remove()`` uninstalls)."""
    handles = []
    for module, index in marks:
        def pre(mod, args, _i=index):
            if not args or not isinstance(args[0], torch.Tensor) or not args[0].requires_grad \
                    or not torch.is_grad_enabled():
                return None
            return (_BucketMark.apply(args[0], callback, _i),) + tuple(args[1:])
        handles.append(module.register_forward_pre_hook(pre))
    return handles


def bucket_candidates(model: nn.Module, offsets) -> list:
    """[(first flat offset, module)] of the model's blocks in registration (= execution) order: the
    top-level children, with ``Sequential`` / ``ModuleList`` containers opened one level (residual
    stages -> blocks, encoder -> layers).  ``offsets`` = ``ParamArena.offsets``."""
    off_of = {id(p): off for p, (off, _, _) in zip(model.parameters(), offsets)}

    def first(m):
        for p in m.parameters():
            return off_of.get(id(p))
        return None

    out = []
    for child in model.children():
        subs = list(child.children()) if isinstance(child, (nn.Sequential, nn.ModuleList)) else []
        for m in (subs or [child]):
            fo = first(m)
            if fo is not None:
                out.append((fo, m))
    # keep only a strictly increasing chain (a module registered out of execution order is skipped)
    chain, last = [], -1
    for fo, m in out:
        if fo > last:
            chain.append((fo, m))
            last = fo
    return chain


# --------------------------------------------------------------------------- switch
def enable_direct_grads(module: nn.Module, *, side_stream: Optional["torch.cuda.Stream"] = None,
                        enabled: bool = True, branch_stream: Optional["torch.cuda.Stream"] = None) -> GradSink:
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
                sink.convs.append(m)        # (the s2d stem also keeps a KRSC shadow for its fallback path)
        elif isinstance(m, FusedBatchNorm2d):
            m._direct_grad = enabled
        if getattr(m, "supports_branch_stream", False):
            # residual blocks run their projection shortcut on this stream (models/resnet.py)
            m._branch_stream = branch_stream if enabled else None
    return sink


__all__ = ["ArenaConv2d", "ArenaLinear", "FusedMaxPool2d", "GradSink", "enable_direct_grads", "S2DStemConv2d",
           "PackedStemInput", "pack_stem_input", "install_bucket_marks", "bucket_candidates"]

"""ResNet-18/34/50 (He et al.) written against plain ``torch.nn``.

The heavy lifting (convolutions, batch-norm) stays in cuDNN/cuBLAS via PyTorch
(model fwd/bwd is library territory); the framework's own kernels
take over at the gradient arena.  Module and parameter names follow the
torchvision layout (``conv1, bn1, layer1.0.conv1, ..., fc``) so a reference
``state_dict`` checkpoint loads with ``strict=True``.
"""
from __future__ import annotations

from typing import List, Sequence, Type, Union

import torch
import torch.nn as nn


def _conv(cin: int, cout: int, k: int, stride: int = 1, padding: int = 0) -> nn.Conv2d:
    """``nn.Conv2d`` subclass that can produce its weight gradient in place in the gradient arena,
    on a side stream (ops/fused_layers.py); identical to ``nn.Conv2d`` until a device worker
    switches that mode on."""
    from ..ops.fused_layers import ArenaConv2d

    return ArenaConv2d(cin, cout, k, stride=stride, padding=padding, bias=False)


def _conv3x3(cin: int, cout: int, stride: int = 1) -> nn.Conv2d:
    return _conv(cin, cout, 3, stride, 1)


def _conv1x1(cin: int, cout: int, stride: int = 1) -> nn.Conv2d:
    return _conv(cin, cout, 1, stride)


def _bn(channels: int, relu: bool = False) -> nn.BatchNorm2d:
    """BatchNorm2d with the following ReLU folded in.  ``FusedBatchNorm2d`` runs the hand-written
    sm_100a NHWC bf16 kernels on CUDA and falls back to ``F.batch_norm`` (+ relu) elsewhere; its
    parameters / buffers / state_dict keys are those of ``nn.BatchNorm2d``."""
    from ..ops.fused_bn import FusedBatchNorm2d

    return FusedBatchNorm2d(channels, relu=relu)


def _shortcut(block: nn.Module, x: torch.Tensor):
    """The block's identity / projection shortcut.  When a device worker has handed the block a
    ``_branch_stream``, the projection (1x1 conv + BN) runs on that stream concurrently with the main
    branch -- forward here, and backward too, because autograd replays an op on the stream of its
    forward.  At one replica per GPU the step is a dependency chain of ~400 short kernels; the three
    projection branches of a ResNet are ~8 % of it."""
    if block.downsample is None:
        return x, None
    side = getattr(block, "_branch_stream", None)
    if side is None or not x.is_cuda:
        return block.downsample(x), None
    main = torch.cuda.current_stream(x.device)
    side.wait_stream(main)
    x.record_stream(side)
    with torch.cuda.stream(side):
        idt = block.downsample(x)
    return idt, (main, side)


def _join(idt: torch.Tensor, fork) -> torch.Tensor:
    if fork is not None:
        main, side = fork
        main.wait_stream(side)
        idt.record_stream(main)
    return idt


class BasicBlock(nn.Module):
    """Two 3 x 3 convolutions with an identity (or 1 x 1 projection) shortcut -- the ResNet-18 / 34 block.  BatchNorm,
    ReLU and the residual add are fused (one kernel per BatchNorm), and the projection shortcut runs on a side stream
    when the model is captured with branch streams.
    """

    expansion = 1
    supports_branch_stream = True

    def __init__(self, cin: int, width: int, stride: int = 1, downsample: nn.Module | None = None):
        super().__init__()
        self.conv1 = _conv3x3(cin, width, stride)
        self.bn1 = _bn(width, relu=True)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = _conv3x3(width, width)
        self.bn2 = _bn(width, relu=True)     # relu(bn2(.) + identity) in one kernel
        self.downsample = downsample

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        idt, fork = _shortcut(self, x)
        y = self.conv2(self.bn1(self.conv1(x)))          # ReLU fused into bn1
        return self.bn2(y, _join(idt, fork))             # residual add + ReLU fused into bn2


class Bottleneck(nn.Module):
    """1 x 1 reduce, 3 x 3, 1 x 1 expand (x 4) with a shortcut -- the ResNet-50 block; same fusions as
    :class:`BasicBlock`.
    """

    expansion = 4
    supports_branch_stream = True

    def __init__(self, cin: int, width: int, stride: int = 1, downsample: nn.Module | None = None):
        super().__init__()
        self.conv1 = _conv1x1(cin, width)
        self.bn1 = _bn(width, relu=True)
        self.conv2 = _conv3x3(width, width, stride)
        self.bn2 = _bn(width, relu=True)
        self.conv3 = _conv1x1(width, width * 4)
        self.bn3 = _bn(width * 4, relu=True)  # relu(bn3(.) + identity) in one kernel
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        idt, fork = _shortcut(self, x)
        y = self.bn1(self.conv1(x))          # ReLU fused
        y = self.conv3(self.bn2(self.conv2(y)))          # ReLU fused
        return self.bn3(y, _join(idt, fork))             # residual add + ReLU fused


class ResNet(nn.Module):
    """ResNet v1 (torchvision layout and parameter names) built from fused layers.

    Parameters
    ----------
    block : BasicBlock or Bottleneck
    depths : sequence of 4 ints
        Blocks per stage.
    num_classes : int, default 1000
    in_channels : int, default 3
    small_input : bool, default False
        CIFAR-style stem (3 x 3 convolution, no max-pool) for 32 x 32 inputs.

    Notes
    -----
    On sm_100a the 7 x 7 / stride-2 stem runs as a space-to-depth 4 x 4 convolution, BatchNorm + ReLU (+ residual) are
    single kernels (one launch per activation through thread-block clusters when it fits, ``csrc/bn.cu``), and weight
    gradients can be written directly into a gradient arena.  On CPU every fused layer falls back to the equivalent
    ``torch.nn`` computation, with the same parameters.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.models import resnet18
    >>> net = resnet18(num_classes=10, small_input=True).eval()
    >>> net(torch.zeros(1, 3, 32, 32)).shape, sum(p.numel() for p in net.parameters())
    (torch.Size([1, 10]), 11173962)
    """

    def __init__(self, block: Type[Union[BasicBlock, Bottleneck]], depths: Sequence[int],
                 num_classes: int = 1000, in_channels: int = 3, small_input: bool = False):
        super().__init__()
        self._cin = 64
        if small_input:  # CIFAR-style stem
            self.conv1 = _conv(in_channels, 64, 3, 1, 1)
            self.maxpool = nn.Identity()
        else:
            from ..ops.fused_layers import FusedMaxPool2d, S2DStemConv2d

            # 7x7/s2 stem: runs as a space-to-depth 4x4 conv on B200 when in_channels == 3
            self.conv1 = S2DStemConv2d(in_channels, 64, 7, stride=2, padding=3, bias=False)
            self.maxpool = FusedMaxPool2d(3, stride=2, padding=1)
        self.bn1 = _bn(64, relu=True)
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = self._stage(block, 64, depths[0], 1)
        self.layer2 = self._stage(block, 128, depths[1], 2)
        self.layer3 = self._stage(block, 256, depths[2], 2)
        self.layer4 = self._stage(block, 512, depths[3], 2)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        from ..ops.fused_layers import ArenaLinear

        self.fc = ArenaLinear(512 * block.expansion, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def _stage(self, block, width: int, depth: int, stride: int) -> nn.Sequential:
        down = None
        if stride != 1 or self._cin != width * block.expansion:
            down = nn.Sequential(_conv1x1(self._cin, width * block.expansion, stride),
                                 _bn(width * block.expansion))
        layers: List[nn.Module] = [block(self._cin, width, stride, down)]
        self._cin = width * block.expansion
        layers += [block(self._cin, width) for _ in range(1, depth)]
        return nn.Sequential(*layers)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.maxpool(self.bn1(self.conv1(x)))   # ReLU fused into bn1
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(torch.flatten(self.avgpool(x), 1))


def resnet18(num_classes: int = 1000, **kw) -> ResNet:
    """ResNet-18 (11.7 M parameters at 1000 classes): the model of the headline benchmark."""
    return ResNet(BasicBlock, (2, 2, 2, 2), num_classes, **kw)


def resnet34(num_classes: int = 1000, **kw) -> ResNet:
    """ResNet-34 (21.8 M parameters at 1000 classes)."""
    return ResNet(BasicBlock, (3, 4, 6, 3), num_classes, **kw)


def resnet50(num_classes: int = 1000, **kw) -> ResNet:
    """ResNet-50 (25.6 M parameters at 1000 classes), bottleneck blocks."""
    return ResNet(Bottleneck, (3, 4, 6, 3), num_classes, **kw)

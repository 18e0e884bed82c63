"""The MNIST ``SmallCNN`` used throughout the reference's examples and its
published ParameterServer benchmark (architecture per reference
examples/ps/nodes.py:46-61: conv(1->32,3) - pool - conv(32->64,3) - pool -
fc(3136->128) - fc(128->10)); parameter names match so checkpoints interchange.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class SmallCNN(nn.Module):
    """The MNIST network of the reference's examples and of its published parameter-server benchmark.

    ``conv(in, 32, 3) - relu - pool - conv(32, 64, 3) - relu - pool - fc(3136, 128) - relu - fc(128, classes)`` for
    28 x 28 inputs; 421 642 parameters with the defaults.  Parameter names match the reference's class, so state dicts
    interchange.

    Parameters
    ----------
    in_channels : int, default 1
    num_classes : int, default 10

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.models import SmallCNN
    >>> net = SmallCNN()
    >>> net(torch.zeros(2, 1, 28, 28)).shape, sum(p.numel() for p in net.parameters())
    (torch.Size([2, 10]), 421642)
    """

    def __init__(self, in_channels: int = 1, num_classes: int = 10):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, 32, 3, padding=1)
        self.conv2 = nn.Conv2d(32, 64, 3, padding=1)
        self.fc1 = nn.Linear(64 * 7 * 7, 128)
        self.fc2 = nn.Linear(128, num_classes)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = F.max_pool2d(F.relu(self.conv1(x)), 2)
        x = F.max_pool2d(F.relu(self.conv2(x)), 2)
        x = F.relu(self.fc1(torch.flatten(x, 1)))
        return self.fc2(x)

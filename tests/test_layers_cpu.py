"""CPU behaviour of the replica layer types: off the fast path they are exactly torch.nn."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from byzpy_b200 import ops
from byzpy_b200.models import resnet18, resnet50
from byzpy_b200.ops.fused_bn import FusedBatchNorm2d
from byzpy_b200.ops.fused_layers import (ArenaConv2d, ArenaLinear, FusedMaxPool2d, S2DStemConv2d,
                                          enable_direct_grads)


def test_resnet18_matches_torchvision_on_cpu_and_loads_its_state_dict():
    torchvision = pytest.importorskip("torchvision")
    torch.manual_seed(0)
    ref = torchvision.models.resnet18(num_classes=7)
    mine = resnet18(num_classes=7)
    mine.load_state_dict(ref.state_dict(), strict=True)
    assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
    x = torch.randn(3, 3, 64, 64)
    for m in (ref, mine):
        m.train()
    ya, yb = ref(x), mine(x)
    torch.testing.assert_close(yb, ya, rtol=1e-5, atol=1e-5)
    ya.sum().backward()
    yb.sum().backward()
    for (n, p), q in zip(ref.named_parameters(), mine.parameters()):
        torch.testing.assert_close(q.grad, p.grad, rtol=1e-4, atol=1e-5, msg=lambda m: f"{n}: {m}")
    # running statistics were updated identically
    torch.testing.assert_close(mine.bn1.running_mean, ref.bn1.running_mean)
    assert int(mine.bn1.num_batches_tracked) == 1


def test_resnet50_forward_shape_and_param_count():
    m = resnet50(num_classes=10).eval()
    assert sum(p.numel() for p in m.parameters()) == 23_528_522
    with torch.no_grad():
        assert m(torch.randn(1, 3, 64, 64)).shape == (1, 10)


def test_layer_fallbacks_equal_torch_nn():
    torch.manual_seed(1)
    x = torch.randn(2, 8, 9, 9)
    conv, ref_conv = ArenaConv2d(8, 4, 3, padding=1, bias=False), nn.Conv2d(8, 4, 3, padding=1, bias=False)
    ref_conv.load_state_dict(conv.state_dict())
    assert torch.equal(conv(x), ref_conv(x))
    stem, ref_stem = S2DStemConv2d(3, 6, 7, stride=2, padding=3, bias=False), nn.Conv2d(3, 6, 7, stride=2, padding=3, bias=False)
    ref_stem.load_state_dict(stem.state_dict())
    img = torch.randn(2, 3, 20, 20)
    assert torch.equal(stem(img), ref_stem(img)) and stem._s2d_ok()
    lin, ref_lin = ArenaLinear(5, 3), nn.Linear(5, 3)
    ref_lin.load_state_dict(lin.state_dict())
    v = torch.randn(4, 5)
    assert torch.equal(lin(v), ref_lin(v))
    assert torch.equal(FusedMaxPool2d(3, stride=2, padding=1)(x), F.max_pool2d(x, 3, 2, 1))
    bn = FusedBatchNorm2d(8, relu=True)
    r = torch.randn_like(x)
    exp = F.relu(F.batch_norm(x, None, None, bn.weight, bn.bias, True, 0.1, 1e-5) + r)
    torch.testing.assert_close(bn(x, r), exp)
    # switching direct gradients on is harmless on CPU: every layer keeps taking the torch path
    net = nn.Sequential(ArenaConv2d(8, 8, 3, padding=1, bias=False), bn)
    conv = net[0]
    sink = enable_direct_grads(net)
    y = net(x)
    y.sum().backward()
    sink.join()
    assert conv.weight.grad is not None and bn.weight.grad is not None
    enable_direct_grads(net, enabled=False)
    assert not conv._direct_grad and not bn._direct_grad


def test_normalize_uint8_nhwc_cpu_fallback():
    x = torch.randint(0, 256, (2, 5, 6, 3), dtype=torch.uint8)
    y = ops.normalize_uint8_nhwc(x, [120.0, 125.0, 130.0], [60.0, 61.0, 62.0])
    assert y.shape == (2, 3, 5, 6) and y.dtype == torch.bfloat16
    m = torch.tensor([120.0, 125.0, 130.0]).view(1, 3, 1, 1)
    s = torch.tensor([60.0, 61.0, 62.0]).view(1, 3, 1, 1)
    torch.testing.assert_close(y.float(), (x.permute(0, 3, 1, 2).float() - m) / s, rtol=1e-2, atol=1e-2)
    z = ops.normalize_uint8_nhwc(x, s2d=True)          # no CUDA: plain layout, same values
    assert z.shape == (2, 3, 5, 6)
    with pytest.raises(TypeError):
        ops.normalize_uint8_nhwc(x.float())
    assert ops.gram_with_median([torch.randn(10) for _ in range(3)]) is None

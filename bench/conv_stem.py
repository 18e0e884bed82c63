"""Micro-benchmark: cuDNN time of the ResNet stem conv (7x7/2, 3->64, batch 32, 224^2, bf16 NHWC)
with the input padded to 3 / 4 / 8 channels, forward and weight gradient."""
import torch

torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
aten = torch.ops.aten


def timeit(fn, it=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3


for cin in (3, 4, 8):
    x = torch.randn(32, cin, 224, 224, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(64, cin, 7, 7, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = aten.convolution(x, w, None, (2, 2), (3, 3), (1, 1), False, (0, 0), 1)
    dy = torch.randn_like(y)
    f = timeit(lambda: aten.convolution(x, w, None, (2, 2), (3, 3), (1, 1), False, (0, 0), 1))
    b = timeit(lambda: aten.convolution_backward(dy, x, w, None, (2, 2), (3, 3), (1, 1), False, (0, 0), 1,
                                                 [False, True, False]))
    print(f"cin={cin}: fwd {f:.1f} us  wgrad {b:.1f} us")

import copy, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from byzpy_b200.models import resnet18
from byzpy_b200.ops.fused_layers import enable_direct_grads
from byzpy_b200.parallel.arena import ParamArena

torch.manual_seed(0)
dev = torch.device("cuda", 0)
base = resnet18(num_classes=10).to(dev)
x = torch.randn(8, 3, 64, 64, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
y = torch.randint(0, 10, (8,), device=dev)
res = {}
for mode in ("autograd", "direct", "direct_nostem"):
    m = copy.deepcopy(base)
    arena = ParamArena(m)
    sink = None
    if mode != "autograd":
        sink = enable_direct_grads(m, side_stream=None)
        if mode == "direct_nostem":
            m.conv1._direct_grad = False
    arena.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = torch.nn.functional.cross_entropy(m(x), y)
    loss.backward()
    if sink is not None:
        sink.join()
    torch.cuda.synchronize()
    res[mode] = ({n: p.grad.clone() for n, p in m.named_parameters()}, loss.item())
print({k: v[1] for k, v in res.items()})
a = res["autograd"][0]
for mode in ("direct", "direct_nostem"):
    b = res[mode][0]
    print("==", mode)
    for n in a:
        rel = ((a[n] - b[n]).norm() / (a[n].norm() + 1e-12)).item()
        if rel > 2e-2:
            print(f"  {n:40s} rel={rel:.3f} |a|={a[n].norm().item():.3e} |b|={b[n].norm().item():.3e}")

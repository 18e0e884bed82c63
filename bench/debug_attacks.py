import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from byzpy_b200.attacks import EmpireAttack, LittleAttack, MimicAttack, SignFlipAttack
DEV = torch.device("cuda", 0)
gen = torch.Generator().manual_seed(6)
g = [torch.randn(4097, generator=gen) for _ in range(7)]
gd = [x.to(DEV) for x in g]
for rep in range(3):
    for name, mk, kw in [("signflip", lambda: SignFlipAttack(scale=-2.0), dict(base_grad=0)),
                   ("empire", lambda: EmpireAttack(scale=-1.1), dict(honest_grads=1)),
                   ("little", lambda: LittleAttack(f=2), dict(honest_grads=1)),
                   ("mimic", lambda: MimicAttack(epsilon=3), dict(honest_grads=1))]:
        kc = {k: (g[0] if v == 0 else g) for k, v in kw.items()}
        kg = {k: (gd[0] if v == 0 else gd) for k, v in kw.items()}
        a, b = mk().apply(**kc), mk().apply(**kg)
        err = (b.cpu() - a).abs().max().item()
        print(rep, name, f"max abs err {err:.3e}")
X = torch.stack(g)
mu = X.mean(0); sd = X.std(0, unbiased=False)
from byzpy_b200 import ops
for a_, b_ in ((1.0, 0.0), (0.0, 1.0), (1.0, 0.3186)):
    out = ops.colstat(gd, a_, b_)
    exp = a_ * mu + b_ * sd
    print("colstat", a_, b_, (out.cpu() - exp).abs().max().item(), (out.cpu().double() - (a_ * X.double().mean(0) + b_ * X.double().std(0, unbiased=False))).abs().max().item())

"""Run-to-run reproducibility of the gradients at the bench shape (debugging aid)."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hcunet_b200 as H
from oracle import unet_oracle as O

torch.manual_seed(0)
m = H.Unet_Constructor(**O.README_3D); m.precision = "mixed"; m = m.cuda().train()
g = torch.Generator().manual_seed(7)
S = (4, 4, 256, 256, 32)
x = torch.randn(S, generator=g).half().cuda()
mask = (torch.rand((4, 1) + S[2:], generator=g) > 0.7).half().cuda()
pwl = (torch.rand((4, 1) + S[2:], generator=g) * 3).half().cuda()
sd = copy.deepcopy(m.state_dict())

def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

runs = []
for i in range(5):
    m.load_state_dict(sd)
    m.zero_grad(set_to_none=True)
    out = m(x)
    loss = H.cross_entropy(out, mask, pwl, "pixel")
    loss.backward()
    torch.cuda.synchronize()
    runs.append(({k: p.grad.detach().clone() for k, p in m.named_parameters()}, out.detach().clone(), float(loss)))
names = [k for k in runs[0][0] if k.endswith("weight") and "batch" not in k]
print("losses", [r[2] for r in runs])
print("logits rel diff vs run 4:", [rel(r[1], runs[4][1]) for r in runs[:4]])
for k in names:
    print(f"{k:28s}", " ".join(f"{rel(runs[i][0][k], runs[4][0][k]):.2e}" for i in range(4)))

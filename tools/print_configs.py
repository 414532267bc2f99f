import os, sys
sys.path.insert(0, "/root/repo")
os.environ["HCU_TC_DEBUG"] = "8"
import torch, hcunet_b200 as H
from oracle import unet_oracle as O
torch.manual_seed(0)
m = H.Unet_Constructor(**O.README_3D); m.precision = "mixed"; m = m.cuda().train()
x = torch.randn(4, 4, 256, 256, 32).cuda().half()
msk = (torch.rand(4, 1, 256, 256, 32) > 0.7).half().cuda(); pwl = torch.rand(4, 1, 256, 256, 32).half().cuda()
for i in range(3):
    if i == 2: print("=== steady step", file=sys.stderr, flush=True)
    out = m(x); loss = H.cross_entropy(out, msk, pwl, "pixel"); loss.backward()
    torch.cuda.synchronize()

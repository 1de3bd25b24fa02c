import os, sys, copy
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from summarizer_b200.models import vasnet_autograd
from summarizer_b200.models.vasnet import VASNet
vasnet_autograd.draw_keep_masks = lambda *a, **k: None
dev = torch.device("cuda")
torch.manual_seed(0)
m = VASNet().to(dev).train()
x = torch.rand(300, 1, 1024, device=dev); x = x / x.norm(dim=2, keepdim=True); tgt = torch.rand(300, 1, 1, device=dev)
crit = torch.nn.MSELoss()
o = torch.optim.Adam(m.parameters(), lr=1e-4, weight_decay=1e-5, fused=True, capturable=True)
def fwd(mod, tag):
    with torch.enable_grad():
        y = mod(x)
    print(f"{tag}: train-forward loss {float(crit(y, tgt)):.5f} range [{float(y.min()):.4f}, {float(y.max()):.4f}]", flush=True)
def step():
    o.zero_grad(set_to_none=True)
    loss = crit(m(x), tgt); loss.backward(); o.step()
    return loss.detach()
print("step loss", float(step()))
fwd(m, "A after step (natural invalidation)")
m._shadow_key = None; fwd(m, "B after manual invalidation")
P = copy.deepcopy(m); P._shadow_key = None; fwd(P, "C deepcopy")
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    z = torch.zeros(4, device=dev) + 1
g.replay(); torch.cuda.synchronize()
fwd(m, "D after an unrelated capture")
m._shadow_key = None; fwd(m, "E after unrelated capture + invalidation")
g2 = torch.cuda.CUDAGraph(); m._shadow_key = None
with torch.cuda.graph(g2):
    with torch.enable_grad():
        y2 = m(x).detach()
g2.replay(); torch.cuda.synchronize()
print(f"F captured train-forward at W1: loss {float(crit(y2, tgt)):.5f}")
m._shadow_key = None; fwd(m, "G eager after that")

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
def snap(): return {n: p.detach().clone() for n, p in m.named_parameters()}
def dmax(a, b): return {n: float((a[n] - b[n]).abs().max()) for n in a if float((a[n] - b[n]).abs().max()) > 0}
def step():
    o.zero_grad(set_to_none=True)
    loss = crit(m(x), tgt); loss.backward(); o.step()
    return loss.detach()
s0 = snap(); l = step(); torch.cuda.synchronize(); s1 = snap()
print("eager step loss", float(l), "moved", max(dmax(s0, s1).values()))
with torch.no_grad():
    m.eval(); print("scores after eager step: range", float(m(x).min()), float(m(x).max())); m.train()
torch.cuda.synchronize(); g = torch.cuda.CUDAGraph(); m._shadow_key = None
with torch.cuda.graph(g):
    out = step()
torch.cuda.synchronize(); s2 = snap()
print("params changed by CAPTURE:", dmax(s1, s2))
g.replay(); torch.cuda.synchronize(); s3 = snap()
print("replay loss", float(out), "moved", max(dmax(s2, s3).values()))
print("moved per param in replay:", {k: round(v, 6) for k, v in dmax(s2, s3).items()})
print("moved per param in eager :", {k: round(v, 6) for k, v in dmax(s0, s1).items()})

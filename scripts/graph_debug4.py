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
def fb():
    for p in m.parameters(): p.grad = None
    y = m(x); loss = crit(y, tgt); loss.backward()
    return y.detach(), loss.detach()
y0, l0 = fb(); g0 = [p.grad.clone() for p in m.parameters()]
# 1. forward only, no grad
m._shadow_key = None
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    with torch.no_grad():
        m.eval(); yi = m(x); m.train()
g.replay(); torch.cuda.synchronize()
with torch.no_grad():
    m.eval(); m._shadow_key = None; yr = m(x); m.train()
print("inference graph vs eager:", float((yi - yr).abs().max()))
# 2. training forward only
m._shadow_key = None
g2 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g2):
    yt = m(x).detach()
g2.replay(); torch.cuda.synchronize()
print("train-forward graph vs eager:", float((yt - y0).abs().max()), "loss", float(crit(yt, tgt)), float(l0))
# 3. forward + backward
m._shadow_key = None
g3 = torch.cuda.CUDAGraph()
for p in m.parameters(): p.grad = None
with torch.cuda.graph(g3):
    y3, l3 = fb()
g3.replay(); torch.cuda.synchronize()
print("fwd+bwd graph: y diff", float((y3 - y0).abs().max()), "loss", float(l3), float(l0),
      "grad diff", max(float((p.grad - q).abs().max()) for p, q in zip(m.parameters(), g0)))
g3.replay(); torch.cuda.synchronize()
print("second replay: y diff", float((y3 - y0).abs().max()), "grad diff", max(float((p.grad - q).abs().max()) for p, q in zip(m.parameters(), g0)))

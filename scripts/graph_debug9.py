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
def fwd(tag):
    with torch.enable_grad():
        y = m(x)
    print(f"{tag}: loss {float(crit(y, tgt).detach()):.5f}", flush=True)
x0 = x.clone()
o.zero_grad(set_to_none=True); loss = crit(m(x), tgt); loss.backward(); o.step(); torch.cuda.synchronize()
s1 = snap(); sh1 = {k: v.clone() for k, v in m._shadow.items()}
fwd("A"); torch.cuda.synchronize(); s2 = snap(); shA = {k: v.clone() for k, v in m._shadow.items()}
print("params changed by forward A:", dmax(s1, s2), "x changed:", float((x - x0).abs().max()))
m._shadow_key = None
fwd("B"); torch.cuda.synchronize(); s3 = snap(); shB = {k: v.clone() for k, v in m._shadow.items()}
print("params changed by forward B:", dmax(s2, s3))
print("shadow A vs B:", {k: float((shA[k].float() - shB[k].float()).abs().max()) for k in shA})
print("shadow A vs W0-shadow:", {k: float((shA[k].float() - sh1[k].float()).abs().max()) for k in shA})
print("key A == natural?", m._shadow_key is not None)

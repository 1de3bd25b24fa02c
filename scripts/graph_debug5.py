import os, sys, copy
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from summarizer_b200.models import vasnet_autograd
from summarizer_b200.models.vasnet import VASNet
vasnet_autograd.draw_keep_masks = lambda *a, **k: None
dev = torch.device("cuda")
torch.manual_seed(0)
base = VASNet().to(dev).train()
x = torch.rand(300, 1, 1024, device=dev); x = x / x.norm(dim=2, keepdim=True); tgt = torch.rand(300, 1, 1, device=dev)
crit = torch.nn.MSELoss()
def make(capturable=True, fused=True):
    m = copy.deepcopy(base); m._shadow_key = None
    return m, torch.optim.Adam(m.parameters(), lr=1e-4, weight_decay=1e-5, fused=fused, capturable=capturable)
def step(m, opt):
    opt.zero_grad(set_to_none=True)
    loss = crit(m(x), tgt); loss.backward(); opt.step()
    return loss.detach()
me, oe = make()
ref = [float(step(me, oe)) for _ in range(4)]
print("eager losses", ref)
for variant in ("warm1", "warm0", "nonfused"):
    m, o = make(fused=(variant != "nonfused"))
    n_warm = 0 if variant == "warm0" else 1
    got = [float(step(m, o)) for _ in range(n_warm)]
    torch.cuda.synchronize(); g = torch.cuda.CUDAGraph(); m._shadow_key = None
    try:
        with torch.cuda.graph(g):
            out = step(m, o)
        for _ in range(4 - n_warm):
            g.replay(); torch.cuda.synchronize(); got.append(float(out)); m._shadow_key = None
        print(variant, got)
    except Exception as e:
        print(variant, "capture failed", type(e).__name__, str(e)[:200])
# which weights does the replayed forward see?  compare the in-graph shadow with the live parameter right after a replay
m, o = make(); step(m, o)
torch.cuda.synchronize(); g = torch.cuda.CUDAGraph(); m._shadow_key = None
with torch.cuda.graph(g):
    out = step(m, o)
sh = m._shadow
before = m.V.weight.detach().clone()
g.replay(); torch.cuda.synchronize()
print("shadow wv vs param BEFORE replay's update:", float((sh["wv"].float() - before.to(torch.bfloat16).float()).abs().max()),
      " vs param AFTER:", float((sh["wv"].float() - m.V.weight.detach().to(torch.bfloat16).float()).abs().max()))
print("b1 alias is param:", sh["b1"].data_ptr() == m.k1.bias.data_ptr())

import os, sys, copy
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from summarizer_b200.models import vasnet_autograd
from summarizer_b200.models.vasnet import VASNet
vasnet_autograd.draw_keep_masks = lambda *a, **k: None
dev = torch.device("cuda")
torch.manual_seed(0)
base = VASNet().to(dev).train()
vids = []
for T in (300, 517):
    x = torch.rand(T, 1, 1024, device=dev); vids.append((x / x.norm(dim=2, keepdim=True), torch.rand(T, 1, 1, device=dev)))
crit = torch.nn.MSELoss()

def make(capturable):
    m = copy.deepcopy(base); m._shadow_key = None
    return m, torch.optim.Adam(m.parameters(), lr=1e-4, weight_decay=1e-5, fused=True, capturable=capturable)

def full_step(m, opt, v):
    opt.zero_grad(set_to_none=True)
    loss = crit(m(vids[v][0]), vids[v][1])
    loss.backward(); opt.step()
    return loss.detach()

def diff(a, b):
    return max(float((p - q).abs().max()) for p, q in zip(a.parameters(), b.parameters()))

me, oe = make(False)
mc, oc = make(True)
mg, og = make(True)
seq = [0, 1, 0, 1, 0, 1]
for v in seq: full_step(me, oe, v)
for v in seq: full_step(mc, oc, v)
print("eager vs eager-capturable:", diff(me, mc))
pool = torch.cuda.graph_pool_handle(); graphs = {}
for i, v in enumerate(seq):
    if i < 2:
        full_step(mg, og, v)
    else:
        if v not in graphs:
            torch.cuda.synchronize(); g = torch.cuda.CUDAGraph(); mg._shadow_key = None
            with torch.cuda.graph(g, pool=pool):
                out = full_step(mg, og, v)
            graphs[v] = (g, out)
        graphs[v][0].replay(); mg._shadow_key = None
    torch.cuda.synchronize()
    # replicate on a fresh eager model up to step i
    mr, orr = make(True)
    for u in seq[:i + 1]: full_step(mr, orr, u)
    print(f"after step {i} (video {v}): graph-path vs eager diff {diff(mg, mr):.3e}")

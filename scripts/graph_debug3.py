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
def make():
    m = copy.deepcopy(base); m._shadow_key = None
    return m, torch.optim.Adam(m.parameters(), lr=1e-4, weight_decay=1e-5, fused=True, capturable=True)
def full_step(m, opt):
    opt.zero_grad(set_to_none=True)
    loss = crit(m(x), tgt)
    loss.backward(); opt.step()
    return loss.detach()
def state(m, opt):
    ps = list(m.parameters())
    return dict(params=[p.detach().clone() for p in ps], grads=[p.grad.detach().clone() for p in ps],
                steps=[float(opt.state[p]["step"]) for p in ps], m=[opt.state[p]["exp_avg"].clone() for p in ps])
def cmp(a, b, tag):
    for k in ("params", "grads", "m"):
        print(f"  {tag} {k}: max diff {max(float((u - v).abs().max()) for u, v in zip(a[k], b[k])):.3e}  (max ref {max(float(v.abs().max()) for v in b[k]):.3e})")
    print(f"  {tag} steps graph {a['steps'][:3]} eager {b['steps'][:3]}")
me, oe = make(); mg, og = make()
full_step(me, oe); full_step(mg, og)
cmp(state(mg, og), state(me, oe), "after eager step 1")
torch.cuda.synchronize(); g = torch.cuda.CUDAGraph(); mg._shadow_key = None
with torch.cuda.graph(g):
    out = full_step(mg, og)
sg = state(mg, og)
print("after capture (not replayed): steps", sg["steps"][:3], "param diff vs eager", max(float((u - v).abs().max()) for u, v in zip(sg["params"], state(me, oe)["params"])))
g.replay(); mg._shadow_key = None; torch.cuda.synchronize()
l2 = full_step(me, oe)
print("loss graph", float(out), "eager", float(l2))
cmp(state(mg, og), state(me, oe), "after step 2 (replay)")
g.replay(); mg._shadow_key = None; torch.cuda.synchronize()
l3 = full_step(me, oe)
print("loss graph", float(out), "eager", float(l3))
cmp(state(mg, og), state(me, oe), "after step 3 (replay)")

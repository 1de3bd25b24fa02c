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
m = copy.deepcopy(base); m._shadow_key = None
o = torch.optim.Adam(m.parameters(), lr=1e-4, weight_decay=1e-5, fused=True, capturable=True)
keep = {}
def step(with_opt=True, zero=True):
    if zero: o.zero_grad(set_to_none=True)
    y = m(x); loss = crit(y, tgt); loss.backward()
    if with_opt: o.step()
    keep["y"] = y.detach(); keep["sh"] = m._shadow
    return loss.detach()
step()
for variant in ("zero+fb+opt", "fb+opt(no zero_grad inside)", "zero+fb (no opt)"):
    torch.cuda.synchronize(); g = torch.cuda.CUDAGraph(); m._shadow_key = None
    if "no zero" in variant: o.zero_grad(set_to_none=True)
    with torch.cuda.graph(g):
        out = step(with_opt="no opt" not in variant, zero="no zero" not in variant)
    P = copy.deepcopy(m); P._shadow_key = None
    g.replay(); torch.cuda.synchronize(); m._shadow_key = None
    with torch.enable_grad():
        yr = P(x); lr_ = crit(yr, tgt)
    sh = keep["sh"]; shr = P._weights(inference=False)[0]
    print(variant, "| loss graph", float(out), "eager", float(lr_), "| scores diff", float((keep["y"] - yr).abs().max()),
          "| score range graph", float(keep["y"].min()), float(keep["y"].max()),
          "| shadow diffs", {k: float((sh[k].float() - shr[k].float()).abs().max()) for k in sh})

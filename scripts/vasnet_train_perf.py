"""VASNet per-video training step: eager launches vs CUDA-graph replay (frames/s, TVSum-like lengths)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from summarizer_b200.models.vasnet import VASNet
dev = torch.device("cuda")
rng = np.random.default_rng(2)
lens = [int(t) for t in rng.integers(167, 1295, size=16)]
g = torch.Generator(device=dev); g.manual_seed(3)
vids = []
for T in lens:
    x = torch.randn(T, 1, 1024, generator=g, device=dev).abs_()
    vids.append((x / x.norm(dim=2, keepdim=True), torch.rand(T, 1, 1, generator=g, device=dev)))
torch.manual_seed(0)
vas = VASNet().to(dev).train()
opt = torch.optim.Adam(vas.parameters(), lr=5e-5, weight_decay=1e-5, fused=True, capturable=True)
crit = torch.nn.MSELoss()

def step(x, t):
    loss = crit(vas(x), t)
    loss.backward(); opt.step()
    return loss

def timed(fn, seconds=3.0):
    for v in range(len(vids)): fn(v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.perf_counter(); e0.record()
    while time.perf_counter() - t0 < seconds:
        for v in range(len(vids)): fn(v)
        n += 1
    e1.record(); torch.cuda.synchronize()
    return n * sum(lens) / (e0.elapsed_time(e1) / 1e3)

def eager(v):
    opt.zero_grad(); step(*vids[v])
out = {"eager_frames_per_s": timed(eager)}
pool = torch.cuda.graph_pool_handle()
graphs = []
for x, t in vids:
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    opt.zero_grad(set_to_none=True)
    with torch.cuda.graph(gr, pool=pool):
        loss = step(x, t)
    graphs.append((gr, loss))
def replay(v):
    graphs[v][0].replay(); vas._shadow_key = None
out["graph_frames_per_s"] = timed(replay)
out["loss_last"] = float(graphs[-1][1])
print(json.dumps(out))

"""DSN REINFORCE step under CUDA-graph replay: full step vs recurrence-only (where does the time go)."""
import json, os, sys, time
import numpy as np, torch
from torch.distributions import Bernoulli
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from summarizer_b200.models.dsn import DSN, compute_rewards
dev = torch.device("cuda")
rng = np.random.default_rng(2)
lens = [int(t) for t in rng.integers(167, 1295, size=16)]
g = torch.Generator(device=dev); g.manual_seed(3)
vids = []
for T in lens:
    x = torch.randn(T, 1, 1024, generator=g, device=dev).abs_()
    vids.append(x / x.norm(dim=2, keepdim=True))
torch.manual_seed(0)
dsn = DSN().to(dev).train()
opt = torch.optim.Adam(dsn.parameters(), lr=5e-5, weight_decay=1e-5, fused=True, capturable=True)
base = torch.zeros((), device=dev)

def full(x):
    opt.zero_grad(set_to_none=True)
    probs = dsn(x)
    dist = Bernoulli(probs, validate_args=False)
    actions = dist.sample((5,))
    rewards = compute_rewards(x, actions.reshape(5, -1))
    loss = -(dist.log_prob(actions).reshape(5, -1).mean(1) * (rewards - base)).sum() / 5.
    loss.backward(); torch.nn.utils.clip_grad_norm_(dsn.parameters(), 5.0); opt.step()

def recurrence_only(x):
    opt.zero_grad(set_to_none=True)
    dsn(x).sum().backward(); opt.step()

def forward_only(x):
    with torch.no_grad():
        dsn(x)

def rate(step):
    for x in vids: step(x)
    pool, graphs = torch.cuda.graph_pool_handle(), []
    for x in vids:
        torch.cuda.synchronize(); gr = torch.cuda.CUDAGraph(); dsn._shadow_key = None
        with torch.cuda.graph(gr, pool=pool):
            step(x)
        graphs.append(gr)
    for gr in graphs: gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        for gr in graphs: gr.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    return dict(ms_per_pass=ms, frames_per_s=sum(lens) / ms * 1e3, us_per_frame=ms * 1e3 / sum(lens))
print(json.dumps(dict(full=rate(full), recurrence_only=rate(recurrence_only), forward_only=rate(forward_only))))

"""DSN scorer timings (CUDA events): recurrence latency per step and batched throughput."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from summarizer_b200.models.dsn import DSN
from summarizer_b200.models.dsn_autograd import dsn_apply

def t(fn, iters=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

torch.manual_seed(0)
m = DSN().cuda().eval()
for nv, T in [(1, 320), (1, 707), (1, 2000), (8, 2000), (64, 2000), (256, 320)]:
    x = torch.rand(nv * T, 1024, device="cuda"); x = x / x.norm(dim=1, keepdim=True)
    ms = t(lambda: m.score_packed(x, [T] * nv))
    print(json.dumps({"mode": "forward", "videos": nv, "T": T, "ms": round(ms, 3), "us_per_step_per_video": round(1e3 * ms / T / max(1, nv / 8), 3),
                      "frames_per_s": round(nv * T / ms * 1e3)}), flush=True)
m.train()
for nv, T in [(1, 707), (8, 707)]:
    x = torch.rand(nv * T, 1024, device="cuda"); x = x / x.norm(dim=1, keepdim=True)
    def step():
        for p in m.parameters(): p.grad = None
        dsn_apply(m, x, [T] * nv).sum().backward()
    ms = t(step)
    print(json.dumps({"mode": "forward+backward", "videos": nv, "T": T, "ms": round(ms, 3), "frames_per_s": round(nv * T / ms * 1e3)}), flush=True)

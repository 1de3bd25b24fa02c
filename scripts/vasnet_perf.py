"""VASNet scoring throughput on sweep-shaped input (bf16 features, T=2000 per video)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from summarizer_b200.models.vasnet import VASNet

def flops_fwd(T, D=1024):
    return 10 * T * D * D + 4 * T * T * D + 2 * T * D

def run(n_videos, T, dtype, iters=5):
    torch.manual_seed(0)
    m = VASNet().cuda().eval()
    x = torch.rand(n_videos * T, 1024, device="cuda")
    x = (x / x.norm(dim=1, keepdim=True)).to(dtype)
    lengths = [T] * n_videos
    for _ in range(2):
        s = m.score_packed(x, lengths, check=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters):
        s = m.score_packed(x, lengths, check=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = n_videos * flops_fwd(T)
    print(json.dumps({"videos": n_videos, "T": T, "dtype": str(dtype), "ms": round(ms, 3),
                      "frames_per_s": round(n_videos * T / ms * 1e3), "tflops": round(fl / ms / 1e9, 1),
                      "finite": bool(torch.isfinite(s).all())}), flush=True)

if __name__ == "__main__":
    run(64, 2000, torch.bfloat16)
    run(256, 2000, torch.bfloat16)
    run(256, 320, torch.float32)
    run(1, 320, torch.float32, iters=20)

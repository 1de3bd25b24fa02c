"""Uninitialised-read detector: poison the caching allocator's free blocks with NaN bit patterns, then run the
training / inference paths; any pad or scratch region that is read before being written shows up as NaN."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from summarizer_b200.models.vasnet import VASNet
from summarizer_b200.models.dsn import DSN
dev = torch.device("cuda")

def poison(mb=3000):
    blocks = [torch.full((mb * 1024 * 1024 // 4 // 8,), float("nan"), device=dev) for _ in range(8)]
    small = [torch.full((n,), float("nan"), device=dev) for n in (256, 1024, 4096, 65536, 1 << 20) for _ in range(8)]
    torch.cuda.synchronize()
    del blocks, small

torch.manual_seed(0)
vas = VASNet().to(dev).train()
dsn = DSN().to(dev).train()
rng = np.random.default_rng(0)
for T in [167, 200, 333, 460, 721, 1000, 1110, 1294, 1301, 64, 65, 10]:
    x = torch.rand(T, 1, 1024, device=dev); x = x / x.norm(dim=2, keepdim=True)
    tgt = torch.rand(T, 1, 1, device=dev)
    for name, m in (("vasnet", vas), ("dsn", dsn)):
        for p in m.parameters(): p.grad = None
        m._shadow_key = None
        poison()
        y = m(x)
        loss = torch.nn.functional.mse_loss(y, tgt)
        loss.backward()
        bad = [n for n, p in m.named_parameters() if p.grad is None or not bool(torch.isfinite(p.grad).all())]
        poison()
        with torch.no_grad():
            m.eval(); yi = m(x); m.train()
        print(f"{name} T={T}: train scores finite={bool(torch.isfinite(y).all())} loss={loss.item():.5f} bad grads={bad} infer finite={bool(torch.isfinite(yi).all())}", flush=True)

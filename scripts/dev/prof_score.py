"""ncu target: two score_packed calls over 32 sweep-shaped videos (2 chunks): the first warms up, the second is profiled."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from summarizer_b200.models.vasnet import VASNet
torch.manual_seed(0)
m = VASNet().cuda().eval()
nv, T = int(os.environ.get("NV", 32)), 2000
x = torch.rand(nv * T, 1024, device="cuda"); x = (x / x.norm(dim=1, keepdim=True)).bfloat16()
for _ in range(2):
    s = m.score_packed(x, [T] * nv, check=False)
torch.cuda.synchronize()
assert m.check_status() and bool(torch.isfinite(s).all())

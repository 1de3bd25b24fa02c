"""SMZ_PROFILE=1 python scripts/vasnet_steps.py — per-step device time of the VASNet scoring stage (warm caches)."""
import os, sys
os.environ["SMZ_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from summarizer_b200 import _native as N
from summarizer_b200.models.vasnet import VASNet
torch.manual_seed(0)
m = VASNet().cuda().eval()
nv, T = 64, 2000
x = torch.rand(nv * T, 1024, device="cuda"); x = (x / x.norm(dim=1, keepdim=True)).bfloat16()
for _ in range(2):
    m.score_packed(x, [T] * nv, check=False)
N.lib().smz_profile_report()          # discard warm-up
for _ in range(5):
    m.score_packed(x, [T] * nv, check=False)
N.lib().smz_profile_report()

"""Relative error of the fast and the exact inference path against the reference goldens (development aid)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.gen_golden_models import VASNET_CASES, build_vasnet, make_input
from summarizer_b200.models.vasnet import VASNet
g = np.load("tests/golden/models_golden.npz")
for name, seed, T, B, kw, sharpen in VASNET_CASES:
    m = build_vasnet(VASNet, seed, kw, sharpen).cuda()
    x = make_input(seed, T, B).cuda()
    if m.max_length is not None:
        continue
    packed = x.permute(1, 0, 2).reshape(B * T, 1024)
    want = torch.from_numpy(g[f"{name}/y"]).cuda().permute(1, 0, 2).reshape(-1)
    for mode in ("fast", "fast-bf16-x", "exact"):
        y = m.score_packed(packed.bfloat16() if "bf16" in mode else packed, [T] * B, exact=(mode == "exact"))
        rel = ((y - want).abs() / want.abs().clamp_min(1e-6))
        print(f"{name:14s} {mode:11s} p50 {rel.median().item():.2e} p95 {torch.quantile(rel, 0.95).item():.2e} max {rel.max().item():.2e}  abs max {(y-want).abs().max().item():.2e}")

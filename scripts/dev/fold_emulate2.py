"""CPU emulation: cost of each bf16 rounding site of the attention half (float64 truth, one site at a time, and fp16 alternatives)."""
import os, sys, math
import numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.gen_golden_models import VASNET_CASES, build_vasnet, make_input
from oracle import ref_import
ns = ref_import.load()
bf = lambda t: t.float().bfloat16().double()
hf = lambda t: t.float().half().double()
idt = lambda t: t
SITES = ["x", "Wqk", "Wv", "QK", "V", "P", "O", "Wo"]
for name, seed, T, B, kw, sharpen in [VASNET_CASES[i] for i in (2, 3, 4)]:
    m = build_vasnet(ns.vasnet.VASNet, seed, kw, sharpen)
    sd = {k: v.double() for k, v in m.state_dict().items()}
    x = make_input(seed, T, B)[:, 0].double()
    def run(r):
        xb = r["x"](x)
        Q = r["QK"](xb @ r["Wqk"](sd["Q.weight"]).t()); K = r["QK"](xb @ r["Wqk"](sd["K.weight"]).t())
        V = r["V"](xb @ r["Wv"](sd["V.weight"]).t())
        e = (Q @ K.t()) * m.scale
        if m.aperture is not None:
            scope = torch.tril(e, diagonal=m.aperture) * torch.triu(e, diagonal=-m.aperture)
            e = e.masked_fill(scope == 0, float("-inf"))
        P = r["P"](torch.exp(e))
        O = r["O"]((P @ V) / torch.exp(e).sum(1, keepdim=True))
        y = O @ r["Wo"](sd["attention_head_projection.weight"]).t() + x
        g, b = sd["layer_norm.weight"], sd["layer_norm.bias"]
        h = torch.relu(F.layer_norm(y, (1024,), g, b, 1e-6) @ sd["k1.weight"].t() + sd["k1.bias"])
        h = F.layer_norm(h, (1024,), g, b, 1e-6)
        return torch.sigmoid(h @ sd["k2.weight"].t() + sd["k2.bias"]).reshape(-1)
    truth = run({s: idt for s in SITES})
    def report(tag, r):
        y = run(r); rel = ((y - truth).abs() / truth.abs())
        print(f"{name:12s} {tag:28s} p50 {rel.median():.2e} p95 {rel.quantile(0.95):.2e} max {rel.max():.2e} bias {((y-truth)/truth).mean():+.2e}")
    for s in SITES:
        report("bf16 only at " + s, {k: (bf if k == s else idt) for k in SITES})
    report("all bf16", {k: bf for k in SITES})
    report("all f16 except P bf16", {k: (bf if k == "P" else hf) for k in SITES})
    report("f16: V O Wv Wo; rest bf16", {k: (hf if k in ("V", "O", "Wv", "Wo") else bf) for k in SITES})
    report("f16: V O Wv Wo x; rest bf16", {k: (hf if k in ("V", "O", "Wv", "Wo", "x") else bf) for k in SITES})
    report("f16: all but P, QK", {k: (bf if k in ("P", "QK") else hf) for k in SITES})

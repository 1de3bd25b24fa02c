"""CPU emulation: which rounding points cost what on the vas_t300 golden (float64 truth vs perturbed variants)."""
import os, sys, math
import numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.gen_golden_models import VASNET_CASES, build_vasnet, make_input
from oracle import ref_import
ns = ref_import.load()
bf = lambda t: t.float().bfloat16().double()
hf = lambda t: t.float().half().double()
for name, seed, T, B, kw, sharpen in VASNET_CASES[2:3]:
    m = build_vasnet(ns.vasnet.VASNet, seed, kw, sharpen)
    sd = {k: v.double() for k, v in m.state_dict().items()}
    x = make_input(seed, T, B)[:, 0].double()
    def run(attn_bf16, ln_mode):
        r = bf if attn_bf16 else (lambda t: t)
        xb = r(x)
        Q, K, V = (r(xb @ r(sd[k + ".weight"]).t()) for k in "QKV")
        e = (Q @ K.t()) * m.scale
        P = r(torch.exp(e)) if attn_bf16 else torch.exp(e)
        O = r((P @ V) / torch.exp(e).sum(1, keepdim=True))
        y = O @ r(sd["attention_head_projection.weight"]).t() + x
        g, b = sd["layer_norm.weight"], sd["layer_norm.bias"]
        W1, b1 = sd["k1.weight"], sd["k1.bias"]
        mu = y.mean(1, keepdim=True); rs = 1 / torch.sqrt(y.var(1, unbiased=False, keepdim=True) + 1e-6)
        if ln_mode == "exact":
            h = ((y - mu) * rs * g + b) @ W1.t() + b1
        elif ln_mode == "bf16":       # separate LayerNorm kernel: bf16 yn, bf16 W1
            h = bf((y - mu) * rs * g + b) @ bf(W1).t() + b1
        elif ln_mode == "fold16":     # folded: f16 y, f16 W1g
            W1g = hf(W1 * g[None, :])
            h = rs * (hf(y) @ W1g.t() - mu * W1g.sum(1)[None, :]) + (W1 @ b + b1)
        h = torch.relu(h)
        h = F.layer_norm(h, (1024,), g, b, 1e-6)
        return torch.sigmoid(h @ sd["k2.weight"].t() + sd["k2.bias"]).reshape(-1)
    truth = run(False, "exact")
    for a in (False, True):
        for ln in ("exact", "bf16", "fold16"):
            y = run(a, ln)
            rel = ((y - truth).abs() / truth.abs())
            print(f"{name} attn_bf16={a!s:5} ln={ln:7s} p50 {rel.median():.2e} p95 {rel.quantile(0.95):.2e} max {rel.max():.2e} mean signed {((y-truth)/truth).mean():+.2e}")

"""CPU emulation (float64 truth) of the folded inference algebra:  logits = (x M) x^T with M = Wq^T Wk,  y = softmax . (x Wvo^T) + x with
Wvo = Wo Wv,  k1 on y with the LayerNorm folded.  Which 16-bit format at which site costs what."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.gen_golden_models import VASNET_CASES, build_vasnet, make_input
from oracle import ref_import
ns = ref_import.load()
bf = lambda t: t.float().bfloat16().double()
hf = lambda t: t.float().half().double()
idt = lambda t: t
for name, seed, T, B, kw, sharpen in [VASNET_CASES[i] for i in (1, 2, 3, 4, 5)]:
    m = build_vasnet(ns.vasnet.VASNet, seed, kw, sharpen)
    sd = {k: v.double() for k, v in m.state_dict().items()}
    x = make_input(seed, T, B)[:, 0].double()
    g, b = sd["layer_norm.weight"], sd["layer_norm.bias"]
    W1, b1 = sd["k1.weight"], sd["k1.bias"]
    def tail(y, folded):
        if folded:
            mu = y.mean(1, keepdim=True); rs = 1 / torch.sqrt(y.var(1, unbiased=False, keepdim=True) + 1e-6)
            W1g = hf(W1 * g[None, :])
            h = rs * (hf(y) @ W1g.t() - mu * W1g.sum(1)[None, :]) + (W1 @ b + b1)
        else:
            h = F.layer_norm(y, (1024,), g, b, 1e-6) @ W1.t() + b1
        h = F.layer_norm(torch.relu(h), (1024,), g, b, 1e-6)
        return torch.sigmoid(h @ sd["k2.weight"].t() + sd["k2.bias"]).reshape(-1)
    def mask(e):
        if m.ignore_self:
            e = e.masked_fill(torch.eye(T, dtype=torch.bool), float("-inf"))
        if m.aperture is not None:
            scope = torch.tril(e, diagonal=m.aperture) * torch.triu(e, diagonal=-m.aperture)
            e = e.masked_fill(scope == 0, float("-inf"))
        return e
    e = mask((x @ sd["Q.weight"].t()) @ (x @ sd["K.weight"].t()).t() * m.scale)
    truth = tail(torch.softmax(e, 1) @ (x @ sd["V.weight"].t()) @ sd["attention_head_projection.weight"].t() + x, False)
    M = (sd["Q.weight"].float().t() @ sd["K.weight"].float()).double()          # float32 products, as the host prepares them
    Wvo = (sd["attention_head_projection.weight"].float() @ sd["V.weight"].float()).double()
    def folded(rx, rM, rG, rWvo, rV, rP):
        xb = rx(x)
        G = rG(xb @ rM(M))            # row i: q_i^T Wk  -> logits_ij = G_i . x_j
        e = mask((G @ xb.t()) * m.scale)
        P = rP(torch.exp(e))
        V = rV(xb @ rWvo(Wvo).t())
        return tail((P @ V) / torch.exp(e).sum(1, keepdim=True) + x, True)
    def report(tag, y):
        rel = ((y - truth).abs() / truth.abs())
        print(f"{name:12s} {tag:34s} p50 {rel.median():.2e} p95 {rel.quantile(0.95):.2e} max {rel.max():.2e} bias {((y-truth)/truth).mean():+.2e}")
    report("folded, all exact but y/W1g f16", folded(idt, idt, idt, idt, idt, idt))
    report("bf16 input: all bf16 (P V' bf16)", folded(bf, bf, bf, bf, bf, bf))
    report("f32 input: x M G Wvo f16, V' P bf16", folded(hf, hf, hf, hf, bf, bf))
    report("bf16 only at Wvo", folded(idt, idt, idt, bf, idt, idt))
    report("bf16 only at V'", folded(idt, idt, idt, idt, bf, idt))
    report("bf16 only at M", folded(idt, bf, idt, idt, idt, idt))
    report("bf16 only at G", folded(idt, idt, bf, idt, idt, idt))

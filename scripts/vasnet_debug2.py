import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.gen_golden_models import VASNET_CASES, build_vasnet, make_input
from summarizer_b200.models.vasnet import VASNet

m = build_vasnet(VASNet, 12, {}, 5.0).cuda()
lengths = [2000] * 6 + [1500, 777]
g = torch.Generator(device="cuda"); g.manual_seed(3)
x = torch.rand(sum(lengths), 1024, generator=g, device="cuda"); x = x / x.norm(dim=1, keepdim=True)
full = m.score_packed(x, lengths); full2 = m.score_packed(x, lengths)
print("determinism full vs full2:", (full - full2).abs().max().item())
o = 0
for i, T in enumerate(lengths):
    a = m.score_packed(x[o:o + T], [T]); a2 = m.score_packed(x[o:o + T], [T])
    d = (full[o:o + T] - a).abs()
    print(f"video {i} T={T} row0={o} lead={o & 7}: max diff {d.max().item():.3e} n>1e-5 {(d > 1e-5).sum().item()} argmax {d.argmax().item()} alone-determinism {(a - a2).abs().max().item():.1e}")
    o += T
# emulate the pipeline in torch with bf16 rounding at the same points, against the fp32 golden
gold = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "models_golden.npz"))
def emulate(m, x):   # x (T,1024) fp32 on cuda
    bf = lambda t: t.bfloat16().float()
    xb = bf(x)
    q = bf(xb @ bf(m.Q.weight).t()); k = bf(xb @ bf(m.K.weight).t()); v = bf(xb @ bf(m.V.weight).t())
    s = (q @ k.t()) * m.scale
    T = x.shape[0]
    idx = torch.arange(T, device=x.device)
    if m.ignore_self: s[idx, idx] = float("-inf")
    if m.aperture is not None:
        s[((idx[:, None] - idx[None, :]).abs() > m.aperture) | (s * s == 0)] = float("-inf")
    p = bf(torch.softmax(s, 1))
    o = bf(p @ v)
    y = o @ bf(m.attention_head_projection.weight).t() + x
    yn = bf(torch.nn.functional.layer_norm(y, (1024,), m.layer_norm.weight, m.layer_norm.bias, m.epsilon))
    h = bf(torch.relu(yn @ bf(m.k1.weight).t() + m.k1.bias))
    hn = torch.nn.functional.layer_norm(h, (1024,), m.layer_norm.weight, m.layer_norm.bias, m.epsilon)
    return torch.sigmoid(hn @ m.k2.weight.t() + m.k2.bias)
for name, seed, T, B, kw, sharpen in VASNET_CASES:
    if "max_length" in kw: continue
    mm = build_vasnet(VASNet, seed, kw, sharpen).cuda()
    xin = make_input(seed, T, B).cuda()
    with torch.no_grad():
        y = mm(xin.clone())
        want = torch.from_numpy(gold[f"{name}/y"]).cuda()
        em = torch.stack([emulate(mm, xin[:, b]) for b in range(B)], 1)
    rel = lambda a, b: ((a - b).abs() / b.abs().clamp_min(1e-6)).max().item()
    print(f"{name}: kernel-vs-golden {rel(y, want):.3e}  emulation-vs-golden {rel(em, want):.3e}  kernel-vs-emulation {rel(y, em):.3e}")

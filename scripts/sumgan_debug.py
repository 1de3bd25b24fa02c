"""Prints error metrics of every SumGAN building block against the CPU oracle (no asserts) — one GPU call shows all."""
import os, sys, time, traceback
import numpy as np
import torch
import torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import models_torch as O
from oracle.gen_golden_models import make_input
from summarizer_b200.dense import gemm, linear
from summarizer_b200.models.lstm_stack import ShadowCache, lstm_stack, lstm_decode

dev = torch.device("cuda")


def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30)), float((a - b).abs().max()), float(b.abs().max())


def report(tag, a, b):
    r, m, s = rel(a, b)
    print(f"  {tag:28s} rel_l2 {r:.3e}  max_abs {m:.3e}  (ref max {s:.3e})", flush=True)


def layer_case(H, bi, T, with_state, inp=1024, layers=2, seed=0):
    print(f"lstm_stack H={H} bi={bi} T={T} state={with_state} layers={layers}", flush=True)
    torch.manual_seed(seed)
    lstm = nn.LSTM(inp, H, num_layers=layers, bidirectional=bi)
    nd = 2 if bi else 1
    x = torch.randn(T, inp) * 0.5
    h0 = torch.randn(layers * nd, H) * 0.3 if with_state else None
    c0 = torch.randn(layers * nd, H) * 0.3 if with_state else None
    wy, wh, wc = torch.randn(T, nd * H), torch.randn(layers * nd, H), torch.randn(layers * nd, H)
    # oracle
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in lstm.named_parameters()}
    xo = x.clone().requires_grad_(True)
    h0o = None if h0 is None else h0.clone().requires_grad_(True)
    c0o = None if c0 is None else c0.clone().requires_grad_(True)
    y, hn, cn = O.lstm_stack(sd, "", xo, layers, bi, h0o, c0o)
    ((y * wy).sum() + (hn * wh).sum() + (cn * wc).sum()).backward()
    # device
    lstm_d = lstm.to(dev)
    cache = ShadowCache()
    xd = x.to(dev).requires_grad_(True)
    h0d = None if h0 is None else h0.to(dev).requires_grad_(True)
    c0d = None if c0 is None else c0.to(dev).requires_grad_(True)
    yd, hnd, cnd = lstm_stack(cache, lstm_d, xd, h0d, c0d)
    ((yd * wy.to(dev)).sum() + (hnd * wh.to(dev)).sum() + (cnd * wc.to(dev)).sum()).backward()
    torch.cuda.synchronize()
    report("y", yd, y); report("h_n", hnd, hn); report("c_n", cnd, cn)
    report("dx", xd.grad, xo.grad)
    if with_state:
        report("dh0", h0d.grad, h0o.grad); report("dc0", c0d.grad, c0o.grad)
    for k, p in lstm_d.named_parameters():
        report("d" + k, p.grad, sd[k].grad)


def decode_case(H, T, seed=0):
    print(f"lstm_decode H={H} T={T}", flush=True)
    torch.manual_seed(seed)
    lstm = nn.LSTM(H, H, num_layers=2)
    with torch.no_grad():
        for p in lstm.parameters():
            p.mul_(2.0)                      # stronger recurrence than the default init
    h = torch.randn(2, H) * 0.5
    c = torch.randn(2, H) * 0.5
    w = torch.randn(T, H)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in lstm.named_parameters()}
    ho, co = h.clone().requires_grad_(True), c.clone().requires_grad_(True)
    x = ho.new_zeros(1, H)
    hh, cc, outs = ho, co, []
    for _ in range(T):
        x, hh, cc = O.lstm_stack(sd, "", x, 2, False, hh, cc)
        outs.append(x)
    top = torch.cat(outs, 0)
    (top * w).sum().backward()
    lstm_d = lstm.to(dev)
    hd, cd = h.to(dev).requires_grad_(True), c.to(dev).requires_grad_(True)
    topd = lstm_decode(ShadowCache(), lstm_d, T, hd, cd)
    (topd * w.to(dev)).sum().backward()
    torch.cuda.synchronize()
    report("top", topd, top); report("dh_init", hd.grad, ho.grad); report("dc_init", cd.grad, co.grad)
    for k, p in lstm_d.named_parameters():
        report("d" + k, p.grad, sd[k].grad)


def linear_case(R, I, Oo):
    print(f"linear rows={R} in={I} out={Oo}", flush=True)
    torch.manual_seed(1)
    x, w, b, g = torch.randn(R, I), torch.randn(Oo, I) * 0.05, torch.randn(Oo), torch.randn(R, Oo)
    xo, wo, bo = (t.clone().requires_grad_(True) for t in (x, w, b))
    ((xo @ wo.t() + bo) * g).sum().backward()
    xd, wd, bd = (t.to(dev).requires_grad_(True) for t in (x, w, b))
    yd = linear(xd, wd, bd)
    (yd * g.to(dev)).sum().backward()
    torch.cuda.synchronize()
    report("y", yd, x @ w.t() + b); report("dx", xd.grad, xo.grad); report("dw", wd.grad, wo.grad); report("db", bd.grad, bo.grad)


def chain_case(name, seed, T):
    from summarizer_b200.models.sumgan import SumGAN
    print(f"chain {name}", flush=True)
    G = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sumgan_golden.npz"))
    torch.manual_seed(seed)
    m = SumGAN().to(dev)
    x = make_input(seed, T, 1).to(dev)
    pr = {k: v.to(dev) for k, v in O.sumgan_probes(seed, T).items()}
    t0 = time.time()
    scores = m(x)
    (mu, logvar), c = m.summarizer.vae.e_lstm(x * scores)
    x_hat = m.summarizer.vae.d_lstm(T, mu, c)
    prob, h_last = m.gan(x_hat)
    loss = (x_hat[:, 0] * pr["x_hat"]).sum() + (mu[:, 0] * pr["mu"]).sum() + (logvar[:, 0] * pr["logvar"]).sum() \
        + (h_last[0] * pr["h_last"]).sum() + prob.sum() + (scores.reshape(-1) * pr["scores"]).sum()
    loss.backward()
    torch.cuda.synchronize()
    print(f"  fwd+bwd {time.time()-t0:.3f}s  loss {loss.item():.6f} vs {float(G[name + '/loss']):.6f}")
    for k, v in (("scores", scores), ("mu", mu), ("logvar", logvar), ("c", c), ("x_hat", x_hat), ("prob", prob), ("h_last", h_last)):
        report(k, v.reshape(-1), torch.from_numpy(G[f"{name}/{k}"]).reshape(-1))
    names = list(G[f"{name}/param_names"]); dig = G[f"{name}/grad_digest"]
    params = dict(m.named_parameters())
    for i, n in enumerate(names):
        g = params[n].grad.detach().double().reshape(-1).cpu()
        k = min(8, g.numel())
        head = float((g[:k] - torch.from_numpy(dig[i, 2:2 + k])).norm() / (np.linalg.norm(dig[i, 2:2 + k]) + 1e-30))
        print(f"  grad {n:46s} abs-sum {float(g.abs().sum()):.4e} vs {dig[i,1]:.4e}  head rel {head:.2e}")


for fn, args in [(linear_case, (2, 2048, 2048)), (linear_case, (37, 2048, 1024)),
                 (layer_case, (1024, False, 1, False, 1024, 1)), (layer_case, (1024, False, 7, True)), (layer_case, (1024, True, 6, False)),
                 (layer_case, (2048, False, 5, True)), (decode_case, (1024, 5)), (decode_case, (2048, 6)),
                 (chain_case, ("sumgan_t9", 11, 9)), (chain_case, ("sumgan_t33", 12, 33))]:
    try:
        fn(*args)
    except Exception:
        traceback.print_exc()
        if "CUDA" in traceback.format_exc():
            break

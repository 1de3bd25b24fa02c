"""SumGAN training step (three phases, sumgan.py:415-480) and selector inference timing on one GPU."""
import json, os, sys, time
import torch
import torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.gen_golden_models import make_input
from summarizer_b200.models.sumgan import SumGAN, SumGANTrainer

dev = torch.device("cuda")
torch.manual_seed(0)
m = SumGAN().to(dev).train()


class _H:  # the few hps fields train_step touches
    lr, weight_decay = 5e-5, 1e-5


t = SumGANTrainer.__new__(SumGANTrainer)
t.model, t.hps, t.sup, t.sigma, t.epoch_noise = m, _H, False, 0.3, 0
t.s_e_optimizer = t._adam(list(m.summarizer.s_lstm.parameters()) + list(m.summarizer.vae.e_lstm.parameters()))
t.d_optimizer = t._adam(m.summarizer.vae.d_lstm.parameters())
t.c_optimizer = t._adam(m.gan.c_lstm.parameters())
t.loss_BCE = nn.BCELoss()
for T in (64, 320):
    x = make_input(1, T, 1).to(dev)
    y = torch.rand(T, 1, 1, device=dev)
    for _ in range(2):
        t.train_step(x, y, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        t.train_step(x, y, 1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    with torch.no_grad():
        m(x); torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            m(x)
        e1.record(); torch.cuda.synchronize()
    print(json.dumps(dict(T=T, train_step_ms=ms, train_frames_per_s=T / ms * 1e3, selector_infer_ms=e0.elapsed_time(e1) / n)), flush=True)
# where the time goes: one phase at a time at T=320 with the profile marks
x = make_input(1, 320, 1).to(dev)
for name, fn in (("selector fwd", lambda: m.summarizer.s_lstm(x)), ("encoder fwd", lambda: m.summarizer.vae.e_lstm(x)),
                 ("discriminator fwd", lambda: m.gan(x))):
    with torch.no_grad():
        fn(); torch.cuda.synchronize(); t0 = time.time(); fn(); torch.cuda.synchronize()
    print(f"{name}: {(time.time()-t0)*1e3:.2f} ms")
h = torch.randn(2, 1, 2048, device=dev) * 0.1
with torch.no_grad():
    m.summarizer.vae.d_lstm(320, h, h); torch.cuda.synchronize(); t0 = time.time(); m.summarizer.vae.d_lstm(320, h, h); torch.cuda.synchronize()
print(f"decoder fwd (320 steps): {(time.time()-t0)*1e3:.2f} ms")
hh = h.clone().requires_grad_(True)
out = m.summarizer.vae.d_lstm(320, hh, h); torch.cuda.synchronize(); t0 = time.time(); out.sum().backward(); torch.cuda.synchronize()
print(f"decoder bwd (320 steps): {(time.time()-t0)*1e3:.2f} ms")

"""SumGAN on the sm_100a LSTM kernels (smz_lstm_seq_* / smz_lstm_decode_* + the tcgen05 GEMM) against the CPU float32
oracle (oracle/models_torch.py, pinned to the reference by tests/test_oracle_sumgan.py) and against golden vectors
produced by the unmodified reference (tests/golden/sumgan_golden.npz).

Tolerance: weights, GEMM operands and the streamed recurrent weights are bfloat16 (fp32 accumulation, fp32 state):
the north-star bar for bf16 is 1e-2; measured relative L2 errors are 1e-3 .. 3.5e-3 per tensor."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import models_torch as O
from oracle.gen_golden_models import make_input

pytestmark = pytest.mark.gpu
TOL = 1e-2
GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "sumgan_golden.npz"))
dev = torch.device("cuda")


def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def clone_params(module):
    return {k: v.detach().clone().requires_grad_(True) for k, v in module.named_parameters()}


@pytest.mark.parametrize("rows,n_in,n_out", [(2, 2048, 2048), (37, 2048, 1024), (1, 1024, 8)])
def test_linear_autograd(rows, n_in, n_out):
    from summarizer_b200.dense import linear
    torch.manual_seed(1)
    x, w, b, g = torch.randn(rows, n_in), torch.randn(n_out, n_in) * 0.05, torch.randn(n_out), torch.randn(rows, n_out)
    xo, wo, bo = (t.clone().requires_grad_(True) for t in (x, w, b))
    ((xo @ wo.t() + bo) * g).sum().backward()
    xd, wd, bd = (t.to(dev).requires_grad_(True) for t in (x, w, b))
    yd = linear(xd, wd, bd)
    (yd * g.to(dev)).sum().backward()
    assert rel(yd, x @ w.t() + b) < TOL and rel(xd.grad, xo.grad) < TOL and rel(wd.grad, wo.grad) < TOL
    assert rel(bd.grad, bo.grad) < 1e-5


@pytest.mark.parametrize("H,bi,T,state,layers", [(1024, False, 1, False, 1), (1024, False, 7, True, 2), (1024, True, 6, False, 2),
                                                  (1024, True, 19, True, 1), (2048, False, 5, True, 2)])
def test_lstm_stack_forward_backward(H, bi, T, state, layers):
    from summarizer_b200.models.lstm_stack import ShadowCache, lstm_stack
    torch.manual_seed(H + T)
    lstm = nn.LSTM(1024, H, num_layers=layers, bidirectional=bi)
    nd = 2 if bi else 1
    x = torch.randn(T, 1024) * 0.5
    h0 = torch.randn(layers * nd, H) * 0.3 if state else None
    c0 = torch.randn(layers * nd, H) * 0.3 if state else None
    wy, wh, wc = torch.randn(T, nd * H), torch.randn(layers * nd, H), torch.randn(layers * nd, H)
    sd = clone_params(lstm)
    xo = x.clone().requires_grad_(True)
    h0o = None if h0 is None else h0.clone().requires_grad_(True)
    c0o = None if c0 is None else c0.clone().requires_grad_(True)
    y, hn, cn = O.lstm_stack(sd, "", xo, layers, bi, h0o, c0o)
    ((y * wy).sum() + (hn * wh).sum() + (cn * wc).sum()).backward()
    lstm_d = lstm.to(dev)
    xd = x.to(dev).requires_grad_(True)
    h0d = None if h0 is None else h0.to(dev).requires_grad_(True)
    c0d = None if c0 is None else c0.to(dev).requires_grad_(True)
    yd, hnd, cnd = lstm_stack(ShadowCache(), lstm_d, xd, h0d, c0d)
    ((yd * wy.to(dev)).sum() + (hnd * wh.to(dev)).sum() + (cnd * wc.to(dev)).sum()).backward()
    assert rel(yd, y) < TOL and rel(hnd, hn) < TOL and rel(cnd, cn) < TOL
    assert rel(xd.grad, xo.grad) < TOL
    if state:
        assert rel(h0d.grad, h0o.grad) < TOL and rel(c0d.grad, c0o.grad) < TOL
    for k, p in lstm_d.named_parameters():
        if sd[k].grad.abs().max() > 0:
            assert rel(p.grad, sd[k].grad) < TOL, k
        else:
            assert p.grad.abs().max() == 0, k
    # inference mode: same outputs, nothing kept for backward
    with torch.no_grad():
        y2, _, _ = lstm_stack(ShadowCache(), lstm_d, x.to(dev), None if h0 is None else h0.to(dev), None if c0 is None else c0.to(dev))
    assert torch.equal(y2, yd.detach())


@pytest.mark.parametrize("H,T", [(1024, 5), (2048, 6), (2048, 1)])
def test_lstm_decode_forward_backward(H, T):
    from summarizer_b200.models.lstm_stack import ShadowCache, lstm_decode
    torch.manual_seed(H + T)
    lstm = nn.LSTM(H, H, num_layers=2)
    with torch.no_grad():
        for p in lstm.parameters():
            p.mul_(2.0)                      # stronger recurrence than the default init
    h, c, w = torch.randn(2, H) * 0.5, torch.randn(2, H) * 0.5, torch.randn(T, H)
    sd = clone_params(lstm)
    ho, co = h.clone().requires_grad_(True), c.clone().requires_grad_(True)
    x, hh, cc, outs = ho.new_zeros(1, H), ho, co, []
    for _ in range(T):                       # sumgan.py:110-112
        x, hh, cc = O.lstm_stack(sd, "", x, 2, False, hh, cc)
        outs.append(x)
    top = torch.cat(outs, 0)
    (top * w).sum().backward()
    lstm_d = lstm.to(dev)
    hd, cd = h.to(dev).requires_grad_(True), c.to(dev).requires_grad_(True)
    topd = lstm_decode(ShadowCache(), lstm_d, T, hd, cd)
    (topd * w.to(dev)).sum().backward()
    assert rel(topd, top) < TOL and rel(hd.grad, ho.grad) < TOL and rel(cd.grad, co.grad) < TOL
    for k, p in lstm_d.named_parameters():
        if sd[k].grad.abs().max() > 0:
            assert rel(p.grad, sd[k].grad) < TOL, k
        else:
            assert p.grad.abs().max() == 0, k


@pytest.fixture(scope="module")
def sumgan_seed11():
    from summarizer_b200.models.sumgan import SumGAN
    torch.manual_seed(11)
    return SumGAN().to(dev)


def run_chain(m, x, pr):
    T = x.shape[0]
    scores = m(x)
    (mu, logvar), c = m.summarizer.vae.e_lstm(x * scores)
    x_hat = m.summarizer.vae.d_lstm(T, mu, c)
    prob, h_last = m.gan(x_hat)
    loss = (x_hat[:, 0] * pr["x_hat"]).sum() + (mu[:, 0] * pr["mu"]).sum() + (logvar[:, 0] * pr["logvar"]).sum() \
        + (h_last[0] * pr["h_last"]).sum() + prob.sum() + (scores.reshape(-1) * pr["scores"]).sum()
    return dict(scores=scores, mu=mu, logvar=logvar, c=c, x_hat=x_hat, prob=prob, h_last=h_last), loss


def test_chain_against_reference_golden(sumgan_seed11):
    """selector -> encoder -> step-wise decoder -> discriminator and one backward pass through all 195 M parameters,
    against what the unmodified reference produced for the same seed and input."""
    name, seed, T = "sumgan_t9", 11, 9
    m = sumgan_seed11.train()
    x = make_input(seed, T, 1).to(dev)
    pr = {k: v.to(dev) for k, v in O.sumgan_probes(seed, T).items()}
    out, loss = run_chain(m, x, pr)
    loss.backward()
    for k, v in out.items():
        assert rel(v, torch.from_numpy(GOLDEN[f"{name}/{k}"])) < TOL, k
    assert np.abs(out["scores"].detach().cpu().numpy().reshape(-1) - GOLDEN[f"{name}/scores"]).max() < 1e-3
    assert loss.item() == pytest.approx(float(GOLDEN[f"{name}/loss"]), rel=5e-3)
    names, dig = list(GOLDEN[f"{name}/param_names"]), GOLDEN[f"{name}/grad_digest"]
    params = dict(m.named_parameters())
    for i, n in enumerate(names):
        g = params[n].grad.detach().double().reshape(-1).cpu()
        assert float(g.abs().sum()) == pytest.approx(dig[i, 1], rel=1e-2), n
        k = min(8, g.numel())
        ref_head = torch.from_numpy(dig[i, 2:2 + k])
        assert float((g[:k] - ref_head).norm()) <= 3e-2 * float(ref_head.norm()) + 3e-2 * dig[i, 1] / g.numel(), n
    m.zero_grad(set_to_none=True)


def test_forward_contract_batch3(sumgan_seed11):
    """the reference's __main__ smoke block (sumgan.py:536-564): batch 3, shapes only — plus batch consistency."""
    m = sumgan_seed11.eval()
    x = make_input(5, 10, 3).to(dev)
    with torch.no_grad():
        x_hat, (mu, logvar), scores = m.summarizer(x)
        probs, h = m.gan(x)
        s = m(x)
        s1 = m(x[:, 1:2])
    assert x_hat.shape == x.shape and scores.shape == (10, 3, 1) and mu.shape == logvar.shape == (2, 3, 2048)
    assert probs.shape == (3, 1) and h.shape == (3, 1024) and s.shape == (10, 3, 1)
    assert torch.equal(s[:, 1:2], s1)
    with pytest.raises(RuntimeError):
        m(x.cpu())


def test_trainer_three_phase_step_and_test(tmp_path):
    """SumGANTrainer: VAE pre-training step, the three adversarial updates, evaluation through the batched device path."""
    from summarizer_b200.utils.config import HParameters
    hps = HParameters()
    hps.log_root, hps.tensorboard = str(tmp_path), False
    hps.load_from_args(dict(model="sumgan", use_cuda="yes", splits_files="summe", log_level="error", epochs=5,
                            extra_params={"pretrain_vae": "1"}))
    t = hps.model_class(hps, hps.splits_files[0]).reset()
    assert type(t).__name__ == "SumGANTrainer" and t.sigma == 0.3 and t.epoch_noise == 1 and t.pretrain_vae == 1
    m = t.model.train()
    key = t._get_train_test_keys(0)[0][0]
    x, y = t._video_tensors(key)
    x, y = x[:48], y[:48]
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    t.s_e_optimizer = t._adam(list(m.summarizer.s_lstm.parameters()) + list(m.summarizer.vae.e_lstm.parameters()))
    t.d_optimizer = t._adam(m.summarizer.vae.d_lstm.parameters())
    t.c_optimizer = t._adam(m.gan.c_lstm.parameters())
    t.loss_BCE = nn.BCELoss()
    for epoch in (0, 1):                                                # with and without discriminator input noise
        out = t.train_step(x, y, epoch)
        vals = torch.stack([out[k].float() for k in ("Lse", "Ld", "Lc", "D_x", "D_x_hat", "D_x_hat_p")])
        assert torch.isfinite(vals).all() and out["scores"].shape == (48, 1, 1)
    moved = {n: float((p.detach() - before[n]).abs().max()) for n, p in m.named_parameters()}
    assert all(v > 0 for v in moved.values()), [n for n, v in moved.items() if v == 0]
    assert all(torch.isfinite(p).all() for p in m.parameters())
    avg_corr, (avg_f, max_f) = t.test(0)
    assert np.isfinite([avg_corr, avg_f, max_f]).all() and 0 <= avg_f <= max_f <= 1


@pytest.mark.parametrize("H,bi,B", [(1024, True, 3), (2048, False, 2), (1024, False, 4)])
def test_sequences_sharing_a_launch_match_separate_launches(H, bi, B):
    """B sequences of equal length go through one recurrence launch per layer (weights streamed once per step for all of
    them): outputs equal the one-sequence launches, parameter gradients equal their sum."""
    from summarizer_b200.models.lstm_stack import ShadowCache, lstm_stack
    torch.manual_seed(B + H)
    lstm = nn.LSTM(1024, H, num_layers=2, bidirectional=bi).to(dev)
    nd, T = (2 if bi else 1), 9
    x = (torch.randn(B, T, 1024, device=dev) * 0.5)
    h0, c0 = torch.randn(2 * nd, B, H, device=dev) * 0.3, torch.randn(2 * nd, B, H, device=dev) * 0.3
    w = torch.randn(B, T, nd * H, device=dev)
    xs, hs = x.clone().requires_grad_(True), h0.clone().requires_grad_(True)
    y, hn, cn = lstm_stack(ShadowCache(), lstm, xs, hs, c0)
    ((y * w).sum() + hn.sum() + cn.sum()).backward()
    g_batched = {k: p.grad.clone() for k, p in lstm.named_parameters()}
    gx, gh = xs.grad.clone(), hs.grad.clone()
    lstm.zero_grad(set_to_none=True)
    for b in range(B):
        xb, hb = x[b].clone().requires_grad_(True), h0[:, b].clone().requires_grad_(True)
        yb, hnb, cnb = lstm_stack(ShadowCache(), lstm, xb, hb, c0[:, b])
        ((yb * w[b]).sum() + hnb.sum() + cnb.sum()).backward()
        assert torch.allclose(yb, y[b], atol=1e-6) and torch.allclose(hnb, hn[:, b], atol=1e-6) and torch.allclose(cnb, cn[:, b], atol=1e-6)
        assert rel(gx[b], xb.grad) < 1e-5 and rel(gh[:, b], hb.grad) < 1e-5
    for k, p in lstm.named_parameters():
        assert rel(g_batched[k], p.grad) < 2e-3, k           # bf16 operands of the summed-over-batch weight-gradient GEMMs


def test_decodes_sharing_a_launch_match_separate_launches():
    from summarizer_b200.models.lstm_stack import ShadowCache, lstm_decode
    torch.manual_seed(9)
    H, T, B = 2048, 7, 2
    lstm = nn.LSTM(H, H, num_layers=2).to(dev)
    h, c, w = torch.randn(2, B, H, device=dev) * 0.5, torch.randn(2, B, H, device=dev) * 0.5, torch.randn(B, T, H, device=dev)
    hs, cs_ = h.clone().requires_grad_(True), c.clone().requires_grad_(True)
    top = lstm_decode(ShadowCache(), lstm, T, hs, cs_)
    (top * w).sum().backward()
    g_batched = {k: p.grad.clone() for k, p in lstm.named_parameters()}
    lstm.zero_grad(set_to_none=True)
    for b in range(B):
        hb, cb = h[:, b].clone().requires_grad_(True), c[:, b].clone().requires_grad_(True)
        tb = lstm_decode(ShadowCache(), lstm, T, hb, cb)
        (tb * w[b]).sum().backward()
        assert torch.allclose(tb, top[b], atol=1e-6)
        assert rel(hs.grad[:, b], hb.grad) < 1e-5 and rel(cs_.grad[:, b], cb.grad) < 1e-5
    for k, p in lstm.named_parameters():
        assert rel(g_batched[k], p.grad) < 2e-3, k


def test_more_sequences_than_one_launch_takes(sumgan_seed11):
    """batch 6 > 4 sequences per launch: processed in groups, same scores as one-by-one."""
    m = sumgan_seed11.eval()
    x = make_input(6, 8, 6).to(dev)
    with torch.no_grad():
        s = m(x)
        ones = torch.cat([m(x[:, b:b + 1]) for b in range(6)], 1)
        (mu, logvar), c = m.summarizer.vae.e_lstm(x)
        x_hat = m.summarizer.vae.d_lstm(8, mu, c)
    assert s.shape == (8, 6, 1) and torch.allclose(s, ones, atol=1e-6)
    assert mu.shape == (2, 6, 2048) and c.shape == (2, 6, 2048) and x_hat.shape == (8, 6, 1024)

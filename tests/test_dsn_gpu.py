"""DSN scorer (smz_dsn_forward: tcgen05 input projection + persistent cluster LSTM) against golden outputs of
the UNMODIFIED reference module (tests/golden/models_golden.npz) and the float32 torch restatement."""
import os

import numpy as np
import pytest
import torch

from oracle import models_torch as MT
from oracle.gen_golden_models import DSN_CASES, build_dsn, checksums, make_input
from summarizer_b200.models.dsn import DSN

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "models_golden.npz"))
REL_P95, REL_MAX = 1e-2, 3e-2      # bf16 W_ih / W_hh, fp32 state and accumulation


@pytest.mark.parametrize("case", DSN_CASES, ids=[c[0] for c in DSN_CASES])
def test_rebuilt_weights_are_the_reference_weights(case):
    name, seed, T, B = case
    np.testing.assert_allclose(checksums(build_dsn(DSN, seed)), GOLDEN[f"{name}/checksum"], rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("case", DSN_CASES, ids=[c[0] for c in DSN_CASES])
def test_forward_matches_reference_golden(case):
    name, seed, T, B = case
    m = build_dsn(DSN, seed).cuda()
    x = make_input(seed, T, B).cuda()
    with torch.no_grad():
        y = m(x)
    assert y.shape == (T, B, 1)
    want = torch.from_numpy(GOLDEN[f"{name}/y"]).cuda()
    rel = ((y - want).abs() / want.abs().clamp_min(1e-6)).flatten()
    assert torch.quantile(rel, 0.95).item() < REL_P95 and rel.max().item() < REL_MAX, \
        f"{name}: relative error p95 {torch.quantile(rel, 0.95).item():.3e} max {rel.max().item():.3e}"


@pytest.mark.gpu
def test_ragged_batch_and_long_sequence():
    torch.manual_seed(5)
    m = DSN().cuda().eval()
    with torch.no_grad():      # larger recurrent weights: the state actually depends on the history
        m.rnn.weight_hh_l0.mul_(3.0); m.rnn.weight_hh_l0_reverse.mul_(3.0)
    lengths = [300, 1, 64, 2000, 129]
    xs = [make_input(200 + i, T, 1)[:, 0].cuda() * 8 for i, T in enumerate(lengths)]
    packed = m.score_packed(torch.cat(xs), lengths)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    o = 0
    for x, T in zip(xs, lengths):
        alone = m.score_packed(x, [T])
        assert torch.equal(packed[o:o + T], alone), f"T={T}"
        if T <= 300:
            want = MT.dsn_forward(sd, x.cpu())
            assert torch.allclose(alone.cpu(), want, rtol=2e-2, atol=2e-3), (T, (alone.cpu() - want).abs().max())
        o += T


@pytest.mark.gpu
@pytest.mark.parametrize("lengths", [[64], [300], [50, 77, 8]])
def test_backward_matches_oracle_autograd(lengths):
    """BPTT on the device (smz_dsn_backward) against torch autograd over the float32 restatement."""
    from summarizer_b200.models.dsn_autograd import dsn_apply
    torch.manual_seed(7)
    m = DSN().cuda().train()
    with torch.no_grad():
        m.rnn.weight_hh_l0.mul_(2.0); m.rnn.weight_hh_l0_reverse.mul_(2.0)
    xs = [make_input(300 + i, T, 1)[:, 0].cuda() * 8 for i, T in enumerate(lengths)]
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    target = torch.rand(sum(lengths), generator=g, device="cuda")
    probs = dsn_apply(m, torch.cat(xs), lengths)
    loss = ((probs - target) ** 2).mean()
    loss.backward()
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    ref = torch.cat([MT.dsn_forward(sd, x) for x in xs])
    ref_loss = ((ref - target) ** 2).mean()
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) / ref_loss.item() < 1e-2
    errs = {}
    for name, p in m.named_parameters():
        errs[name] = (p.grad - sd[name].grad).norm().item() / max(sd[name].grad.norm().item(), 1e-12)
    report = ", ".join(f"{k} {v:.2e}" for k, v in errs.items())
    print("relative gradient errors:", report)
    assert max(errs.values()) < 3e-2, report


@pytest.mark.gpu
def test_module_forward_backward_through_nn_module():
    torch.manual_seed(0)
    m = DSN().cuda().train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    x = make_input(9, 120, 1).cuda()
    target = torch.linspace(0.1, 0.9, 120, device="cuda").view(120, 1, 1)
    losses = []
    for _ in range(25):
        loss = torch.nn.functional.mse_loss(m(x), target)
        opt.zero_grad(); loss.backward(); torch.nn.utils.clip_grad_norm_(m.parameters(), 5.0); opt.step()
        losses.append(float(loss))
    assert np.isfinite(losses).all() and losses[-1] < losses[0]


@pytest.mark.gpu
@pytest.mark.parametrize("T,far", [(64, False), (300, False), (707, False), (200, True)])
def test_reward_matches_oracle(T, far):
    """smz_dsn_reward (Gram once per video, all episodes in one pass) vs the float32 restatement of
    compute_reward, episode by episode."""
    from summarizer_b200.models.dsn import compute_rewards
    seq = make_input(500 + T, T, 1)[:, 0]
    g = torch.Generator().manual_seed(T)
    probs = [0.5, 0.3, 0.7, 0.02, 0.0]
    actions = torch.stack([(torch.rand(T, generator=g) < p).float() for p in probs])
    actions[3] = 0; actions[3, T // 2] = 1                      # exactly one picked frame; row 4: none
    got = compute_rewards(seq.cuda(), actions.cuda(), far_sim=far, temp_dist_thre=20).cpu()
    want = torch.tensor([MT.dsn_reward(seq, actions[e], far_sim=far, temp_dist_thre=20) for e in range(len(probs))])
    assert torch.allclose(got, want, rtol=2e-4, atol=1e-6), (got, want)


@pytest.mark.gpu
def test_dsn_trainer_reinforce_runs(tmp_path):
    from summarizer_b200.main import train
    from summarizer_b200.utils.config import HParameters
    hps = HParameters()
    hps.log_root, hps.tensorboard = str(tmp_path), False
    hps.load_from_args(dict(model="dsn", use_cuda="yes", splits_files="splits/summe_splits_overfit.json", log_level="error",
                            epochs=2, test_every_epochs=1, extra_params={}))
    (res,) = train(hps)
    assert np.isfinite(res[1:]).all() and 0 <= res[2] <= res[3] <= 1

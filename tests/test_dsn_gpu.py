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
# bf16 W_ih / W_hh, fp32 state and accumulation: the worst frame of every golden is within 2.7e-5 of the reference
# (the gates saturate little at the reference initialisation), so the scorer is held to BASELINE's FLOAT32 bar, 1e-4
REL_P95, REL_MAX = 5e-5, 1e-4


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
    print(f"{name}: relative error p95 {torch.quantile(rel, 0.95).item():.2e} max {rel.max().item():.2e}")
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


@pytest.mark.gpu
def test_sample_episodes_logprob_and_gradient_match_torch():
    """smz_bernoulli_logprob on GIVEN actions = torch's Bernoulli(probs).log_prob(actions).mean over the frames
    (dsn.py:112,126,135), and its backward = torch's autograd through it, including probabilities at the clamp."""
    from torch.distributions import Bernoulli
    from summarizer_b200.models.dsn import episode_state, sample_episodes
    T, E = 1003, 5
    g = torch.Generator().manual_seed(4)
    probs = torch.rand(T, 1, 1, generator=g)
    probs[:4, 0, 0] = torch.tensor([1e-9, 1 - 1e-7, 0.5, 0.999])
    actions = (torch.rand(E, T, generator=g) < 0.4).float()
    actions[:, 0] = 1; actions[:, 1] = 0                       # improbable actions: steep log-probabilities
    w = torch.randn(E, generator=g)

    p_ref = probs.clone().cuda().requires_grad_(True)
    lp_ref = Bernoulli(p_ref, validate_args=False).log_prob(actions.cuda().reshape(E, T, 1, 1)).reshape(E, -1).mean(1)
    (lp_ref * w.cuda()).sum().backward()

    p = probs.clone().cuda().requires_grad_(True)
    state = episode_state(p.device, seed=7)
    lp, act = sample_episodes(p, E, state, given=actions.cuda())
    (lp * w.cuda()).sum().backward()
    assert act.dtype == torch.uint8 and torch.equal(act.float().cpu(), actions)
    assert state.tolist() == [7, 0, 0]                          # given actions: no draw consumed
    assert torch.allclose(lp, lp_ref, rtol=1e-5, atol=1e-6), (lp, lp_ref)
    assert torch.allclose(p.grad, p_ref.grad, rtol=1e-4, atol=1e-7), (p.grad - p_ref.grad).abs().max()


@pytest.mark.gpu
def test_sample_episodes_draws_are_bernoulli_and_advance_on_the_device():
    """Drawn episodes: P(action = 1) = p per frame, independent between episodes and calls; the call number is bumped
    on the device (a captured step draws new episodes at every replay); the same seed reproduces the draws."""
    from summarizer_b200.models.dsn import episode_state, sample_episodes
    T, E = 4099, 8
    probs = torch.linspace(0.02, 0.98, T).cuda()
    state = episode_state(probs.device, seed=123)
    draws = []
    for call in range(16):
        lp, act = sample_episodes(probs, E, state)
        assert state.tolist() == [123, call + 1, 0]
        ref = torch.where(act.bool(), probs.log(), torch.log1p(-probs)).mean(1)
        assert torch.allclose(lp, ref, rtol=1e-5, atol=1e-6)
        draws.append(act)
    allv = torch.stack(draws).reshape(-1, T).float()            # 128 independent draws per frame
    # mean over frames of (action - p) ~ N(0, sum p(1-p)) / n
    z = (allv - probs).sum() / torch.sqrt((probs * (1 - probs)).sum() * allv.shape[0])
    assert abs(float(z)) < 4.0, float(z)
    by_bucket = (allv.mean(0).reshape(-1)[:4096].reshape(16, 256).mean(1) - probs[:4096].reshape(16, 256).mean(1)).abs()
    assert float(by_bucket.max()) < 0.01, by_bucket
    rows = allv.reshape(-1, T)
    assert len({bytes(r.to(torch.uint8).cpu().numpy().tobytes()) for r in rows}) == rows.shape[0]   # no repeated episode
    again = episode_state(probs.device, seed=123)
    _, act0 = sample_episodes(probs, E, again)
    assert torch.equal(act0, draws[0])
    other = episode_state(probs.device, seed=124)
    assert not torch.equal(sample_episodes(probs, E, other)[1], draws[0])

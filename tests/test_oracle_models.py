"""CPU: the torch restatement of the scorers (oracle/models_torch.py) reproduces the golden outputs of
the UNMODIFIED reference modules (tests/golden/models_golden.npz), and — when /root/reference is present —
the live reference modules."""
import os

import numpy as np
import pytest
import torch

from oracle import models_torch as MT
from oracle.gen_golden_models import DSN_CASES, VASNET_CASES, build_vasnet, make_input
from summarizer_b200.models.vasnet import VASNet

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "models_golden.npz"))


@pytest.mark.parametrize("case", [c for c in VASNET_CASES if "max_length" not in c[4]], ids=lambda c: c[0])
def test_vasnet_restatement_matches_golden(case):
    name, seed, T, B, kw, sharpen = case
    m = build_vasnet(VASNet, seed, kw, sharpen)
    sd = m.state_dict()
    x = make_input(seed, T, B)
    with torch.no_grad():
        y = torch.stack([MT.vasnet_forward(sd, x[:, b], scale=m.scale, eps=m.epsilon, aperture=m.aperture,
                                           ignore_self=m.ignore_self) for b in range(B)], 1)
    np.testing.assert_allclose(y.numpy(), GOLDEN[f"{name}/y"][:, :, 0], rtol=2e-5, atol=2e-6)


@pytest.mark.reference
@pytest.mark.parametrize("case", DSN_CASES[:3], ids=lambda c: c[0])
def test_dsn_restatement_matches_golden(case):
    from oracle import ref_import
    name, seed, T, B = case
    torch.manual_seed(seed)
    ref = ref_import.load().dsn.DSN().eval()
    sd = ref.state_dict()
    x = make_input(seed, T, B)
    with torch.no_grad():
        y = torch.stack([MT.dsn_forward(sd, x[:, b]) for b in range(B)], 1)
    np.testing.assert_allclose(y.numpy(), GOLDEN[f"{name}/y"][:, :, 0], rtol=2e-5, atol=2e-6)


@pytest.mark.reference
def test_dsn_reward_restatement_matches_live_reference():
    """oracle.models_torch.dsn_reward == the reference's DSNTrainer.compute_reward (dsn.py:185-236)."""
    import types
    from oracle import ref_import
    ref = ref_import.load().dsn.DSNTrainer
    me = types.SimpleNamespace(hps=types.SimpleNamespace(use_cuda=False))
    g = torch.Generator().manual_seed(4)
    for T, p, far in [(60, 0.5, False), (200, 0.3, False), (150, 0.6, True), (80, 0.0, False)]:
        seq = make_input(70 + T, T, 1)
        actions = (torch.rand(T, 1, 1, generator=g) < p).float()
        want = float(ref.compute_reward(me, seq, actions, far_sim=far, temp_dist_thre=20))
        got = MT.dsn_reward(seq[:, 0], actions.reshape(-1), far_sim=far, temp_dist_thre=20)
        assert got == pytest.approx(want, rel=1e-5, abs=1e-7)

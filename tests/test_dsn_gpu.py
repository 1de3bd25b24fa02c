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

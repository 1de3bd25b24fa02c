"""CPU: the SumGAN restatement (oracle/models_torch.py) and the parameter layout of
summarizer_b200.models.sumgan.SumGAN against golden vectors produced by the unmodified reference
(oracle/gen_golden_sumgan.py -> tests/golden/sumgan_golden.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle.gen_golden_models import make_input
from oracle.models_torch import sumgan_chain, sumgan_probes

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "sumgan_golden.npz"))


@pytest.fixture(scope="module")
def model_t9():
    from summarizer_b200.models.sumgan import SumGAN
    torch.manual_seed(11)
    return SumGAN()


def test_state_dict_layout_and_init_stream(model_t9):
    names = sorted(n for n, _ in model_t9.named_parameters())
    assert names == list(GOLDEN["sumgan_t9/param_names"])          # reference .pth files load by key
    params = dict(model_t9.named_parameters())
    mine = np.asarray([float(params[n].detach().double().abs().sum()) for n in names])
    np.testing.assert_allclose(mine, GOLDEN["sumgan_t9/checksum"], rtol=1e-12)
    assert sum(p.numel() for p in model_t9.parameters()) == 195_158_018


def test_oracle_chain_matches_reference(model_t9):
    name, seed, T = "sumgan_t9", 11, 9
    sd = dict(model_t9.named_parameters())
    x = make_input(seed, T, 1)[:, 0]
    r = sumgan_chain(sd, x, sumgan_probes(seed, T))
    for k in ("scores", "mu", "logvar", "c", "x_hat", "prob", "h_last"):
        np.testing.assert_allclose(r[k].detach().numpy().reshape(-1), GOLDEN[f"{name}/{k}"].reshape(-1), rtol=2e-4, atol=2e-6, err_msg=k)
    assert r["loss"].item() == pytest.approx(float(GOLDEN[f"{name}/loss"]), rel=1e-5)
    r["loss"].backward()
    names = list(GOLDEN[f"{name}/param_names"])
    dig = GOLDEN[f"{name}/grad_digest"]
    for i, n in enumerate(names):
        g = sd[n].grad.detach().double().reshape(-1)
        scale = dig[i, 1] / max(g.numel(), 1) + 1e-12             # mean |g| of the tensor
        assert abs(float(g.abs().sum()) - dig[i, 1]) <= 2e-3 * dig[i, 1] + 1e-9, n
        k = min(8, g.numel())
        np.testing.assert_allclose(g[:k].numpy(), dig[i, 2:2 + k], rtol=5e-3, atol=5e-3 * scale, err_msg=n)
    for p in model_t9.parameters():
        p.grad = None

"""KTS on the device (smz_kts_gram / smz_kts) against the float64 restatement of the published algorithm
(oracle/kts_np.py).  The DP is compared on the SAME kernel matrix (the device's float32 Gram matrix), so change points
must be identical and the objective equal to float64 rounding."""
import numpy as np
import pytest
import torch

from oracle import kts_np
from summarizer_b200.utils import kts

pytestmark = pytest.mark.gpu


def features(n, d, seed, jumps, noise=0.2):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32) * np.float32(noise)
    level = np.zeros(d, dtype=np.float32)
    for i in range(n):
        if i in jumps:
            level = rng.standard_normal(d).astype(np.float32)
        x[i] += level
    return x


def test_gram_matrix():
    x = features(333, 1024, 0, (50, 200))
    K = kts.gram(x).cpu().numpy()
    ref = x.astype(np.float64) @ x.astype(np.float64).T
    assert np.array_equal(K, K.T)
    np.testing.assert_allclose(K, ref, rtol=2e-5, atol=2e-4)


@pytest.mark.parametrize("n,ncp,lmin,lmax", [(97, 5, 1, 100000), (200, 12, 3, 60), (64, 0, 1, 100000), (150, 7, 1, 40)])
def test_cpd_nonlin_matches_oracle(n, ncp, lmin, lmax):
    x = features(n, 64, n, (n // 5, n // 2, 3 * n // 4))
    K = kts.gram(x)
    cps, scores = kts.cpd_nonlin(K, ncp, lmin=lmin, lmax=lmax)
    rc, rs = kts_np.cpd_nonlin(K.cpu().numpy().astype(np.float64), ncp, lmin=lmin, lmax=lmax)
    assert cps.tolist() == rc.tolist()
    np.testing.assert_allclose(scores, rs, rtol=1e-9, atol=1e-9)


def test_cpd_auto_and_end_to_end():
    jumps = (40, 95, 160, 230)
    x = features(300, 1024, 7, jumps, noise=0.02)      # low noise: the vmax = 1 penalty stops at the planted changes
    K = kts.gram(x)
    cps, scores = kts.cpd_auto(K, 30, vmax=1.0)
    rc, rs = kts_np.cpd_auto(K.cpu().numpy().astype(np.float64), 30, 1.0)
    assert cps.tolist() == rc.tolist() == list(jumps)
    np.testing.assert_allclose(scores, rs, rtol=1e-9, atol=1e-9)
    assert kts.kts(torch.from_numpy(x), 30, vmax=1.0).tolist() == list(jumps)
    noisy = kts.gram(features(300, 1024, 7, jumps))    # high noise: both pick the maximum number of change points
    c2, _ = kts.cpd_auto(noisy, 30, vmax=1.0)
    r2, _ = kts_np.cpd_auto(noisy.cpu().numpy().astype(np.float64), 30, 1.0)
    assert c2.tolist() == r2.tolist() and len(c2) == 30
    cp, nf = kts.segments_from_change_points(cps, 300 * 15, rate=15)
    assert cp[0, 0] == 0 and cp[-1, 1] == 300 * 15 - 1 and nf.sum() == 300 * 15 and len(nf) == len(jumps) + 1


def test_argument_errors():
    K = kts.gram(features(20, 8, 1, ()))
    with pytest.raises(ValueError):
        kts.cpd_nonlin(K, 25)

"""CPU: the KTS restatement (oracle/kts_np.py, published algorithm; parity unpinned — the reference ships no KTS) against
brute force on tiny inputs, and the host-side segment conversion."""
import numpy as np
import pytest

from oracle import kts_np


def make_K(n, d, seed, jumps=()):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)) * 0.1
    level = np.zeros(d)
    for i in range(n):
        if i in jumps:
            level = rng.standard_normal(d)
        x[i] += level
    return x @ x.T


@pytest.mark.parametrize("n,ncp,lmin,lmax", [(9, 2, 1, 100000), (10, 3, 1, 100000), (11, 2, 2, 6), (8, 0, 1, 100000)])
def test_dp_is_optimal(n, ncp, lmin, lmax):
    K = make_K(n, 5, n + ncp, jumps=(3, 6))
    cps, scores = kts_np.cpd_nonlin(K, ncp, lmin=lmin, lmax=lmax)
    best, arg = kts_np.brute_force(K, ncp, lmin, lmax)
    assert scores[ncp] == pytest.approx(best, rel=1e-10, abs=1e-10)
    if ncp:
        assert tuple(cps) == tuple(arg)


def test_cpd_auto_finds_the_planted_changes():
    K = make_K(60, 16, 3, jumps=(17, 41))
    cps, scores = kts_np.cpd_auto(K, 8, vmax=1.0)
    assert list(cps) == [17, 41] and len(scores) == 3 and np.all(np.diff(scores) < 0)


def test_segment_conversion():
    from summarizer_b200.utils.kts import segments_from_change_points, uniform_segments
    cp, nf = segments_from_change_points([4, 10], 200, rate=15)
    assert cp.tolist() == [[0, 59], [60, 149], [150, 199]] and nf.tolist() == [60, 90, 50] and nf.sum() == 200
    cp, nf = uniform_segments(130, 60)
    assert cp.tolist() == [[0, 59], [60, 119], [120, 129]] and nf.tolist() == [60, 60, 10]

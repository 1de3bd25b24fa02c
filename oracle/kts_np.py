"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Kernel temporal segmentation, restated from the PUBLISHED algorithm: D. Potapov, M. Douze, Z. Harchaoui, C. Schmid,
"Category-specific video summarization", ECCV 2014, and the authors' public release (cpd_nonlin.py / cpd_auto.py), which
the video-summarisation datasets used to produce their ``/change_points`` (reference datasets/README.md:24-27).

PARITY UNPINNED: the reference repository ships no KTS code and no KTS fixtures (it only consumes change points), and the
authors' package is not in this image.  What is pinned here: brute-force optimality of the DP on small inputs and the
penalty formula as published.  float64 throughout."""
import numpy as np


def calc_scatters(K):
    """scatters[i, j] = within-segment scatter of frames i..j (inclusive) for the kernel matrix K."""
    K = np.asarray(K, dtype=np.float64)
    n = K.shape[0]
    K1 = np.cumsum([0.0] + list(np.diag(K)))
    K2 = np.zeros((n + 1, n + 1))
    K2[1:, 1:] = np.cumsum(np.cumsum(K, 0), 1)
    scatters = np.zeros((n, n))
    for i in range(n):
        for j in range(i, n):
            scatters[i, j] = K1[j + 1] - K1[i] - (K2[j + 1, j + 1] + K2[i, i] - K2[j + 1, i] - K2[i, j + 1]) / (j - i + 1)
    return scatters


def cpd_nonlin(K, ncp, lmin=1, lmax=100000, backtrack=True):
    """Change points minimising the total within-segment scatter with exactly ``ncp`` change points.
    Returns (cps ascending int array [ncp], scores [ncp+1] = optimal objective for 0..ncp change points)."""
    m = int(ncp)
    n = K.shape[0]
    assert n >= (m + 1) * lmin and n <= (m + 1) * lmax and 1 <= lmin <= lmax
    J = calc_scatters(K)
    I = 1e101 * np.ones((m + 1, n + 1))
    I[0, lmin:lmax] = J[0, lmin - 1:lmax - 1]
    p = np.zeros((m + 1, n + 1), dtype=int)
    for k in range(1, m + 1):
        for l in range((k + 1) * lmin, n + 1):
            tmin = max(k * lmin, l - lmax)
            tmax = l - lmin + 1
            c = J[tmin:tmax, l - 1].reshape(-1) + I[k - 1, tmin:tmax].reshape(-1)
            I[k, l] = np.min(c)
            p[k, l] = np.argmin(c) + tmin
    cps = np.zeros(m, dtype=int)
    if backtrack:
        cur = n
        for k in range(m, 0, -1):
            cps[k - 1] = p[k, cur]
            cur = cps[k - 1]
    scores = I[:, n].copy()
    scores[scores > 1e99] = np.inf
    return cps, scores


def cpd_auto(K, ncp, vmax, desc_rate=1, **kwargs):
    """Number of change points chosen by the penalised objective, then cpd_nonlin with that number."""
    m = int(ncp)
    _, scores = cpd_nonlin(K, m, backtrack=False, **kwargs)
    N = K.shape[0]
    N2 = N * desc_rate
    penalties = np.zeros(m + 1)
    ncp_ = np.arange(1, m + 1)
    penalties[1:] = (vmax * ncp_ / (2.0 * N2)) * (np.log(float(N2) / ncp_) + 1)
    costs = scores / float(N) + penalties
    m_best = int(np.argmin(costs))
    return cpd_nonlin(K, m_best, **kwargs)


def brute_force(K, ncp, lmin=1, lmax=100000):
    """All placements of ncp change points (tiny n only): (best objective, best cps)."""
    from itertools import combinations
    n = K.shape[0]
    J = calc_scatters(K)
    best, arg = np.inf, None
    for cps in combinations(range(1, n), ncp):
        b = [0] + list(cps) + [n]
        if any(not (lmin <= b[i + 1] - b[i] <= lmax) for i in range(len(b) - 1)):
            continue
        v = sum(J[b[i], b[i + 1] - 1] for i in range(len(b) - 1))
        if v < best:
            best, arg = v, cps
    return best, arg

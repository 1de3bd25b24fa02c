"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — ctypes access to oracle/smz_oracle.c."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libsmz_oracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libsmz_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.smzo_mean_f32.restype = C.c_float
        L.smzo_mean_f32.argtypes = [C.c_void_p, C.c_int64]
        L.smzo_generate_summary.restype = C.c_int64
        L.smzo_generate_summary.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                            C.c_int, C.c_void_p, C.c_int64, C.c_double, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.smzo_knapsack.restype = C.c_int
        L.smzo_knapsack.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]
        L.smzo_evaluate_summary.restype = C.c_int
        L.smzo_evaluate_summary.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int64, C.c_int64,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def mean_f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return np.float32(lib().smzo_mean_f32(_p(a), a.shape[0]))


def knapsack(profits, weights, capacity):
    p = np.ascontiguousarray(profits, dtype=np.int64)
    w = np.ascontiguousarray(weights, dtype=np.int64)
    out = np.zeros(len(p), dtype=np.uint8)
    rc = lib().smzo_knapsack(_p(p), _p(w), len(p), int(capacity), _p(out))
    assert rc == 0
    return np.nonzero(out)[0].tolist()


def generate_summary(scores, cps, n_frames, nfps, positions, proportion=0.15, method="knapsack"):
    """Returns (summary, parts) like oracle.eval_np.generate_summary(return_parts=True)."""
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    pos = np.ascontiguousarray(positions, dtype=np.int32)
    cps = np.ascontiguousarray(cps, dtype=np.int32)
    nfps = np.ascontiguousarray(nfps, dtype=np.int32)
    n_segs = cps.shape[0]
    seg = np.zeros(n_segs, dtype=np.float32)
    vals = np.zeros(n_segs, dtype=np.int64)
    picked = np.zeros(n_segs, dtype=np.uint8)
    summary = np.zeros(int(nfps.sum()), dtype=np.float32)
    cap = lib().smzo_generate_summary(_p(scores), len(scores), _p(pos), len(pos), _p(cps), n_segs, _p(nfps),
                                      int(n_frames), float(proportion), {"knapsack": 0, "rank": 1}[method],
                                      _p(seg), _p(vals), _p(picked), _p(summary))
    if cap < 0:
        raise RuntimeError(f"smzo_generate_summary failed ({cap})")
    return summary, dict(seg_score=seg, values=vals, capacity=int(cap), picks=np.nonzero(picked)[0].tolist())


def evaluate_summary(machine_summary, user_summary):
    """Returns dict(overlap, gsum, msum, f, avg_f, max_f) — float32 F (no-padding dtype)."""
    m = np.ascontiguousarray(machine_summary, dtype=np.float32)
    u = np.ascontiguousarray(user_summary, dtype=np.float32)
    n_users, n_frames = u.shape
    ov = np.zeros(n_users, dtype=np.int32)
    gs = np.zeros(n_users, dtype=np.int32)
    ms = np.zeros(1, dtype=np.int32)
    f = np.zeros(n_users, dtype=np.float32)
    avg = np.zeros(1, dtype=np.float64)
    mx = np.zeros(1, dtype=np.float64)
    rc = lib().smzo_evaluate_summary(_p(m), len(m), _p(u), n_users, n_frames, n_frames, _p(ov), _p(gs), _p(ms),
                                     _p(f), _p(avg), _p(mx))
    assert rc == 0
    return dict(overlap=ov, gsum=gs, msum=int(ms[0]), f=f, avg_f=avg[0], max_f=mx[0])

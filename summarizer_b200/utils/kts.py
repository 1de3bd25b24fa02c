"""Kernel temporal segmentation on the device (csrc/smz_kts.cu) with the interface of the authors' public KTS release
(``cpd_nonlin`` / ``cpd_auto``), plus the conversion to the dataset fields ``change_points`` / ``n_frame_per_seg``
(reference datasets/README.md:24-30) and the uniform 2-second segmentation the paper uses for Twitch-LOL.

The reference consumes precomputed change points and ships no KTS code (SURVEY.md §8f NEXT-4); this is the step right
before the hot path.  No CPU fallback: the functions need a CUDA (sm_100a) device."""
import ctypes as C

import numpy as np
import torch

from .. import _native as N


def _run(K, features, n, ncp, lmin, lmax, auto, vmax, desc_rate):
    N.require_device()
    m = int(ncp)
    if not (n >= (m + 1) * lmin and n <= (m + 1) * lmax and 1 <= lmin <= lmax):
        raise ValueError(f"need (ncp+1)*lmin <= n <= (ncp+1)*lmax (n={n}, ncp={m}, lmin={lmin}, lmax={lmax})")
    lmax = int(min(lmax, n + 1))
    dev = (K if K is not None else features).device
    nbytes = C.c_int64(0)
    N.check(N.lib().smz_kts_workspace_bytes(n, m, int(K is None), C.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    cps = torch.zeros(max(m, 1), dtype=torch.int32, device=dev)
    n_cps = torch.zeros(1, dtype=torch.int32, device=dev)
    scores = torch.empty(m + 1, dtype=torch.float64, device=dev)
    d, ld = (features.shape[1], features.stride(0)) if features is not None else (0, 0)
    N.check(N.lib().smz_kts(N.ptr(K), N.ptr(features), n, d, ld, m, int(lmin), lmax, int(auto), float(vmax), int(desc_rate),
                            N.ptr(cps), N.ptr(n_cps), N.ptr(scores), N.ptr(ws), ws.numel(), N.current_stream()))
    k = int(n_cps.item())
    return cps[:k].cpu().numpy().astype(int), scores.cpu().numpy()


def _device_matrix(K):
    if isinstance(K, np.ndarray):
        K = torch.from_numpy(np.ascontiguousarray(K, dtype=np.float32))
    if not K.is_cuda:
        K = K.cuda()
    return K.float().contiguous()


def gram(features):
    """K = X X^T in float32 on the device: features (n, d) -> (n, n) cuda tensor."""
    N.require_device()
    x = _device_matrix(features)
    n, d = x.shape
    K = torch.empty(n, n, dtype=torch.float32, device=x.device)
    N.check(N.lib().smz_kts_gram(N.ptr(x), n, d, x.stride(0), N.ptr(K), N.current_stream()))
    return K


def cpd_nonlin(K, ncp, lmin=1, lmax=100000, backtrack=True):
    """Exactly ``ncp`` change points minimising the within-segment scatter.  K: (n, n) kernel matrix (numpy or torch).
    Returns (cps int array [ncp] ascending, scores [ncp+1]); cps is all zeros with ``backtrack=False`` (as upstream)."""
    K = _device_matrix(K)
    cps, scores = _run(K, None, K.shape[0], ncp, lmin, lmax, False, 0.0, 1)
    return (cps if backtrack else np.zeros(int(ncp), dtype=int)), scores


def cpd_auto(K, ncp, vmax, desc_rate=1, **kwargs):
    """Number of change points (<= ncp) chosen by the penalised objective; returns (cps, scores[0..len(cps)])."""
    K = _device_matrix(K)
    cps, scores = _run(K, None, K.shape[0], ncp, kwargs.get("lmin", 1), kwargs.get("lmax", 100000), True, vmax, desc_rate)
    return cps, scores[:len(cps) + 1]


def kts(features, max_ncp, vmax=1.0, lmin=1, lmax=100000, desc_rate=1):
    """Features (n, d) -> change points, Gram matrix included (one call, everything on the device)."""
    x = _device_matrix(features)
    cps, _ = _run(None, x, x.shape[0], max_ncp, lmin, lmax, True, vmax, desc_rate)
    return cps


def segments_from_change_points(cps, n_frames, rate=1):
    """Change points in subsampled-frame units (``rate`` original frames per sample, 15 in the datasets) -> the dataset
    fields: change_points (num_segments, 2) inclusive [start, end] in original frames, n_frame_per_seg (num_segments,)."""
    b = np.concatenate([[0], np.asarray(cps, dtype=np.int64) * int(rate), [int(n_frames)]])
    b = np.unique(np.clip(b, 0, int(n_frames)))
    change_points = np.stack([b[:-1], b[1:] - 1], 1).astype(np.int32)
    return change_points, (change_points[:, 1] - change_points[:, 0] + 1).astype(np.int32)


def uniform_segments(n_frames, seg_frames):
    """Fixed-length segmentation (the paper's 2-second shots for Twitch-LOL): same output fields."""
    return segments_from_change_points(np.arange(seg_frames, n_frames, seg_frames), n_frames, 1)

"""Rank correlation on the device (utils/eval.py:49-72): smz_rank_correlation over one video or a batch."""
import numpy as np
import torch

from . import _native as N


class CorrBatch:
    """Resident annotator score rows (dataset /user_scores) of a set of videos + work buffers.
    ``videos``: list of (n_frames, user_scores (n_users, n_frames))."""

    def __init__(self, videos, device=None):
        N.require_device()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        B = len(videos)
        desc = np.zeros(B, dtype=N.CORR_DESC)
        rows, mo, uo, ro, r0 = [], 0, 0, 0, 0
        for i, (n_frames, us) in enumerate(videos):
            us = np.ascontiguousarray(np.asarray(us, dtype=np.float32).reshape(-1, int(n_frames)))
            d = desc[i]
            d["m_off"], d["u_off"], d["u_ld"], d["rank_off"] = mo, uo, int(n_frames), ro
            d["n_frames"], d["n_users"], d["row0"] = int(n_frames), us.shape[0], r0
            rows.append(us.reshape(-1))
            mo += int(n_frames); uo += us.size; ro += (us.shape[0] + 1) * int(n_frames); r0 += us.shape[0]
        self.h_desc, self.n_videos = desc, B
        self.total_frames, self.total_rows = mo, r0
        self.max_n_frames = int(desc["n_frames"].max()) if B else 0
        self.max_n_users = int(desc["n_users"].max()) if B else 0
        dev = self.device
        self.d_desc = torch.from_numpy(desc.view(np.uint8).reshape(-1).copy()).to(dev)
        self.d_user = torch.from_numpy(np.concatenate(rows) if rows else np.zeros(1, np.float32)).to(dev)
        self.rank_ws = torch.empty(max(ro, 1), dtype=torch.float32, device=dev)
        self.corr = torch.empty(max(r0, 1), dtype=torch.float64, device=dev)
        self.corr_avg = torch.empty(max(B, 1), dtype=torch.float64, device=dev)

    def correlate(self, machine_frame_scores, metric="spearmanr"):
        """machine_frame_scores: packed float32 device tensor (sum n_frames,) -> per-video mean correlation."""
        if metric not in N.SMZ_METRIC:
            raise KeyError(f"Unknown metric {metric}")
        m = machine_frame_scores.to(device=self.device, dtype=torch.float32).contiguous()
        if m.numel() < self.total_frames:
            raise ValueError(f"expected {self.total_frames} machine frame scores, got {m.numel()}")
        N.check(N.lib().smz_rank_correlation(N.ptr(self.d_desc), self.n_videos, self.max_n_frames, self.max_n_users,
                                             N.ptr(m), N.ptr(self.d_user), N.SMZ_METRIC[metric], N.ptr(self.rank_ws),
                                             N.ptr(self.corr), N.ptr(self.corr_avg), N.current_stream()))
        return self.corr_avg[: self.n_videos]


def rank_correlation(machine_scores, user_scores, metric="spearmanr"):
    """evaluate_scores for one video: numpy in, numpy float64 out."""
    if metric not in N.SMZ_METRIC:
        raise KeyError(f"Unknown metric {metric}")
    machine = np.ascontiguousarray(np.asarray(machine_scores, dtype=np.float32).reshape(-1))
    user = np.asarray(user_scores)
    n_users, n_frames = user.shape
    if machine.size != n_frames:
        raise ValueError("all the input array dimensions must match")
    b = CorrBatch([(n_frames, user)])
    return np.float64(b.correlate(torch.from_numpy(machine), metric)[0].item())

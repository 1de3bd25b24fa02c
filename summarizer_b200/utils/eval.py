"""Summary generation and evaluation with the reference's signatures (utils/eval.py), computed
by the sm_100a kernels of libsummarizer_b200.so.

Per-video functions take and return numpy arrays exactly like the reference
(utils/eval.py:15,37,49,74,125); each call packs a batch of one video, ships it to the device and
reads the result back.  The batched, device-resident entry points live in
``summarizer_b200.batch.VideoBatch`` (select / fscore over thousands of videos per launch) and are
what ``Trainer.test`` and ``bench.py`` use.
"""
import numpy as np
import torch

from .. import _native as N
from ..batch import VideoBatch


def _single_video_batch(n_frames, positions, cps=None, nfps=None, user_summary=None, n_scores=None,
                        proportion=0.15):
    v = dict(n_frames=int(n_frames), picks=np.asarray(positions).reshape(-1),
             change_points=np.zeros((0, 2), np.int32) if cps is None else cps,
             n_frame_per_seg=np.zeros(0, np.int32) if nfps is None else nfps)
    if n_scores is not None:
        v["n_scores"] = n_scores
    if user_summary is not None:
        v["user_summary"] = user_summary
    return VideoBatch([v], proportion=proportion)


def upsample(scores, n_frames, positions):
    """Upsample scores vector to the original number of frames (utils/eval.py:15-35).
    Input
      scores: (n_steps,)
      n_frames: (1,)
      positions: (n_steps, 1)
    Output
      frame_scores: (n_frames,) float32
    """
    scores = np.ascontiguousarray(np.asarray(scores).reshape(-1), dtype=np.float32)
    n_frames = int(n_frames)
    if n_frames == 0:
        return np.zeros(0, dtype=np.float32)
    b = _single_video_batch(n_frames, positions, n_scores=scores.size)
    out = b.upsample(torch.from_numpy(scores))
    b.check_status()
    return out[:n_frames].cpu().numpy()


def generate_scores(probs, n_frames, positions):
    """Set score to every original frame of the video (utils/eval.py:37-47)."""
    return upsample(probs, n_frames, positions)


def evaluate_scores(machine_scores, user_scores, metric="spearmanr"):
    """Rank correlation between machine and user scores (utils/eval.py:49-72).
    Input
      machine_scores: (n_frames,)
      user_scores: (n_users, n_frames)
    Output
      avg_corr: mean over annotators
    """
    from ..rankcorr import rank_correlation
    if metric not in ("spearmanr", "kendalltau"):
        raise KeyError(f"Unknown metric {metric}")
    return rank_correlation(machine_scores, user_scores, metric)


def generate_summary(scores, cps, n_frames, nfps, positions, proportion=0.15, method="knapsack"):
    """Generate keyshot-based video summary i.e. a binary vector (utils/eval.py:74-123).
    Input
      scores: predicted importance scores (n_steps,)
      cps: change points, (n_segs, 2), inclusive [start, end]
      n_frames: original number of frames
      nfps: number of frames per segment
      positions: positions of subsampled frames in the original video
      proportion: length of video summary (compared to original video length)
      method: 'knapsack' or 'rank'
    Output
      summary: float32 0/1 vector of length sum(nfps)
    """
    if method not in N.SMZ_METHOD:
        raise KeyError(f"Unknown method {method}")
    scores = np.ascontiguousarray(np.asarray(scores).reshape(-1), dtype=np.float32)
    cps = np.asarray(cps)
    b = _single_video_batch(n_frames, positions, cps=cps, nfps=nfps, n_scores=scores.size, proportion=proportion)
    b.select(torch.from_numpy(scores), method=method)
    b.check_status()
    return b.summary_of(0).cpu().numpy()


def evaluate_summary(machine_summary, user_summary):
    """Compare machine summary with user summary (keyshot-based) (utils/eval.py:125-165).
    Input
      machine_summary: (n_frames,)
      user_summary: (n_users, n_frames)
    Output
      avg_f_score, max_f_score
    """
    machine = np.ascontiguousarray(np.asarray(machine_summary).reshape(-1), dtype=np.float32)
    user = np.asarray(user_summary)
    n_users, n_frames = user.shape
    desc = np.zeros(1, dtype=N.VIDEO_DESC)
    ld = (n_frames + 3) // 4 * 4
    desc["user_ld"], desc["n_users"], desc["n_frames"], desc["summ_len"] = ld, n_users, n_frames, machine.size
    buf = np.zeros((n_users, ld), dtype=np.float32)
    buf[:, :n_frames] = user
    empty = np.zeros(0, np.int32)
    b = VideoBatch.from_packed(desc, empty, empty, empty, buf.reshape(-1), proportion=0.15)
    b.pack_summary(torch.from_numpy(machine) if machine.size else torch.zeros(1))
    b.fscore()
    overlap = b.overlap[:n_users].cpu().numpy()
    avg_f, max_f = b.avg_f[0].item(), b.max_f[0].item()
    if machine.size < n_frames:       # zero-padded with float64 zeros (utils/eval.py:143-145): the kernel's float64 branch
        return np.float64(avg_f), np.float64(max_f)
    if (overlap == 0).any():          # the reference's list then holds a Python 0. -> float64 results
        return np.float64(avg_f), np.float64(max_f)
    return np.float32(avg_f), np.float32(max_f)

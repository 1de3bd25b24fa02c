"""Device-resident ragged batches of videos for the shot-selection / F-score kernels.

A ``VideoBatch`` packs the dataset fields the evaluation path reads for a set of videos
(datasets/README.md:5-42: /picks, /change_points, /n_frame_per_seg, /n_frames, /user_summary)
into flat device arrays plus one ``smz_video_desc`` per video, uploads them ONCE and keeps
them resident in HBM; every ``Trainer.test`` call then only ships the model's scores.
The reference re-reads the same fields from HDF5 on every call (models/__init__.py:88-119).
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _native as N


def _round_up(x, m):
    return (x + m - 1) // m * m


def capacity_of(n_frames, proportion):
    """utils/eval.py:96 — int(math.floor(n_frames * proportion)), float64."""
    return int(math.floor(int(n_frames) * float(proportion)))


class VideoBatch:
    """Packed batch.  ``videos`` is a list of dicts with keys
    ``n_frames`` (int), ``picks`` (int array), ``change_points`` ((n_segs,2) int, inclusive ends),
    ``n_frame_per_seg`` (int array) and optionally ``user_summary`` ((n_users,n_frames) 0/1) and
    ``n_scores`` (defaults to len(picks))."""

    def __init__(self, videos, proportion=0.15, device=None, pad_user_rows=True):
        N.require_device()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.proportion = float(proportion)
        B = len(videos)
        desc = np.zeros(B, dtype=N.VIDEO_DESC)
        picks, cps, nfps, users = [], [], [], []
        so = po = sg = uo = sm = fo = mo = uc = 0
        for i, v in enumerate(videos):
            n_frames = int(v["n_frames"])
            p = np.asarray(v["picks"])
            if p.dtype != np.int64:                      # utils/eval.py:25-26
                p = p.astype(np.int32)
            p = np.ascontiguousarray(p, dtype=np.int32)
            c = np.ascontiguousarray(np.asarray(v["change_points"]).reshape(-1, 2), dtype=np.int32)
            w = np.ascontiguousarray(np.asarray(v["n_frame_per_seg"]).reshape(-1), dtype=np.int32)
            if c.shape[0] != w.shape[0]:
                raise ValueError(f"video {i}: {c.shape[0]} change points but {w.shape[0]} n_frame_per_seg")
            if p.size == 0:
                raise ValueError(f"video {i}: empty picks")
            if np.any(np.diff(p) < 0) or p[0] < 0:
                raise ValueError(f"video {i}: picks must be non-negative and ascending")
            if c.size and (np.any(c[:, 0] < 0) or np.any(c[:, 0] > c[:, 1]) or np.any(c[:, 0] >= n_frames)):
                raise ValueError(f"video {i}: change points must satisfy 0 <= start <= end, start < n_frames")
            if np.any(w < 0):
                raise ValueError(f"video {i}: negative n_frame_per_seg")
            n_scores = int(v.get("n_scores", p.size))
            us = v.get("user_summary")
            n_users = 0
            ld = 0
            if us is not None:
                us = np.asarray(us)
                if us.ndim != 2 or us.shape[1] != n_frames:
                    raise ValueError(f"video {i}: user_summary must be (n_users, n_frames={n_frames})")
                n_users = us.shape[0]
                if n_users > N.FSCORE_MAX_USERS:
                    raise ValueError(f"video {i}: more than {N.FSCORE_MAX_USERS} annotators")
                ld = _round_up(n_frames, 4) if pad_user_rows else n_frames
                buf = np.zeros((n_users, ld), dtype=np.float32)
                buf[:, :n_frames] = us
                users.append(buf.reshape(-1))
            d = desc[i]
            d["score_off"], d["picks_off"], d["seg_off"], d["user_off"] = so, po, sg, uo
            d["user_ld"], d["summ_off"], d["frame_off"], d["mask_off"], d["ucount_off"] = ld, sm, fo, mo, uc
            d["n_scores"], d["n_picks"], d["n_segs"], d["n_users"] = n_scores, p.size, w.size, n_users
            d["n_frames"], d["summ_len"] = n_frames, int(w.sum())
            d["capacity"] = capacity_of(n_frames, proportion)
            picks.append(p); cps.append(c.reshape(-1)); nfps.append(w)
            so += n_scores; po += p.size; sg += w.size
            uo += _round_up(n_users * ld, 4)
            if us is not None and (n_users * ld) % 4:
                users.append(np.zeros(4 - (n_users * ld) % 4, dtype=np.float32))
            sm += int(w.sum()); fo += n_frames; mo += (n_frames + 31) // 32; uc += n_users
        cat = lambda xs, dt: np.concatenate(xs).astype(dt, copy=False) if xs else np.zeros(0, dt)
        self._finish(desc, cat(picks, np.int32), cat(cps, np.int32), cat(nfps, np.int32),
                     cat(users, np.float32) if users else None)

    @classmethod
    def from_packed(cls, desc, picks, cps, nfps, user_summary, proportion, device=None):
        """Batch over arrays that are already packed (``user_summary`` may be a device tensor)."""
        N.require_device()
        self = cls.__new__(cls)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.proportion = float(proportion)
        desc = np.ascontiguousarray(desc, dtype=N.VIDEO_DESC)
        if len(desc) and int(desc["n_users"].max()) > N.FSCORE_MAX_USERS:      # fscore_kernel counts in shared memory
            raise ValueError(f"more than {N.FSCORE_MAX_USERS} annotators in a video")
        self._finish(desc, picks, cps, nfps, user_summary)
        return self

    def _to_dev(self, a, dtype):
        if a is None:
            return None
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.device, dtype=dtype)

    def _finish(self, desc, picks, cps, nfps, users):
        self.h_desc = desc
        B = self.n_videos = len(desc)
        self.total_scores = int(desc["n_scores"].sum())
        self.total_segs = int(desc["n_segs"].sum())
        self.total_users = int(desc["n_users"].sum())
        self.total_summary = int(desc["summ_len"].sum())
        self.total_frames = int(desc["n_frames"].sum())
        self.total_mask_words = int(((desc["n_frames"] + 31) // 32).sum())
        self.max_n_segs = int(desc["n_segs"].max()) if B else 0
        self.max_capacity = max(int(desc["capacity"].max()), 0) if B else 0
        self.max_n_frames = int(desc["n_frames"].max()) if B else 0
        if isinstance(nfps, torch.Tensor):
            self.max_seg_frames = int(nfps.max().item()) if nfps.numel() else 0
        else:
            self.max_seg_frames = int(np.max(nfps)) if len(nfps) else 0
        self.has_users = users is not None and self.total_users > 0
        dev = self.device
        self.d_desc = torch.from_numpy(desc.view(np.uint8).reshape(-1).copy()).to(dev)
        self.d_picks = self._to_dev(picks, torch.int32)
        self.d_cps = self._to_dev(cps, torch.int32)
        self.d_nfps = self._to_dev(nfps, torch.int32)
        self.d_users = self._to_dev(users, torch.float32)
        i32 = dict(dtype=torch.int32, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        # outputs (resident, reused by every call)
        self.seg_mean = torch.empty(max(self.total_segs, 1), **f32)
        self.values = torch.empty(max(self.total_segs, 1), **i32)
        self.picked = torch.empty(max(self.total_segs, 1), dtype=torch.uint8, device=dev)
        self.summary = torch.empty(max(self.total_summary, 1), **f32)
        self.mask = torch.empty(max(self.total_mask_words, 1), **i32)
        self.msum = torch.empty(max(B, 1), **i32)
        self.status = torch.zeros(max(B, 1), **i32)
        self.overlap = torch.empty(max(self.total_users, 1), **i32)
        self.gsum = torch.empty(max(self.total_users, 1), **i32)
        self.f = torch.empty(max(self.total_users, 1), **f32)
        self.avg_f = torch.empty(max(B, 1), dtype=torch.float64, device=dev)
        self.max_f = torch.empty(max(B, 1), dtype=torch.float64, device=dev)
        nbytes = ctypes.c_int64(0)
        N.check(N.lib().smz_select_workspace_bytes(B, self.max_n_segs, self.max_capacity, self.max_n_frames,
                                                   self.max_seg_frames, ctypes.byref(nbytes)))
        self.ws_bytes = int(nbytes.value)
        self.ws = torch.empty(max(self.ws_bytes, 4), dtype=torch.uint8, device=dev)

    # ---------------------------------------------------------------------------------------
    def select(self, scores, method="knapsack", write_summary=True):
        """generate_summary for every video (utils/eval.py:74-123).  ``scores``: packed float32
        device tensor (sum n_scores,).  Results land in self.picked / summary / mask / msum."""
        if method not in N.SMZ_METHOD:
            raise KeyError(f"Unknown method {method}")
        scores = self._to_dev(scores, torch.float32)
        if scores.numel() != self.total_scores:
            raise ValueError(f"expected {self.total_scores} scores, got {scores.numel()}")
        N.check(N.lib().smz_select_shots(
            N.ptr(self.d_desc), self.n_videos, N.ptr(scores), N.ptr(self.d_picks), N.ptr(self.d_cps),
            N.ptr(self.d_nfps), N.SMZ_METHOD[method], self.max_n_segs, self.max_capacity, self.max_n_frames,
            self.max_seg_frames, N.ptr(self.seg_mean), N.ptr(self.values), N.ptr(self.picked),
            N.ptr(self.summary) if write_summary else None, N.ptr(self.mask), N.ptr(self.msum),
            N.ptr(self.status), N.ptr(self.ws), self.ws_bytes, N.current_stream()))
        return self

    def evaluate(self, scores, method="knapsack", write_summary=True, d_bits=None):
        """select() followed by fscore() (or fscore_packed(d_bits)) in ONE library call (smz_eval_batch): the persistent
        CTA that solves a video's knapsack also builds its summary mask in shared memory and streams its annotator rows
        against it, so the shared-memory-bound DP of some videos overlaps the HBM-bound F-score of others.  Same
        results as the two separate calls."""
        if method not in N.SMZ_METHOD:
            raise KeyError(f"Unknown method {method}")
        if d_bits is None and not self.has_users:
            raise ValueError("this batch holds no user_summary")
        scores = self._to_dev(scores, torch.float32)
        if scores.numel() != self.total_scores:
            raise ValueError(f"expected {self.total_scores} scores, got {scores.numel()}")
        bits_off = None
        if d_bits is not None:
            self._bits_layout()
            if d_bits.numel() < self.total_bit_words or d_bits.dtype != torch.int32 or not d_bits.is_cuda:
                raise ValueError("evaluate: int32 device tensor of total_bit_words expected")
            bits_off = self.d_bits_off
        N.check(N.lib().smz_eval_batch(
            N.ptr(self.d_desc), self.n_videos, self.total_users,
            N.ptr(scores), N.ptr(self.d_picks), N.ptr(self.d_cps), N.ptr(self.d_nfps), N.SMZ_METHOD[method],
            self.max_n_segs, self.max_capacity, self.max_n_frames, self.max_seg_frames,
            N.ptr(self.d_users) if d_bits is None else None, N.ptr(d_bits), N.ptr(bits_off),
            N.ptr(self.seg_mean), N.ptr(self.values), N.ptr(self.picked),
            N.ptr(self.summary) if write_summary else None, N.ptr(self.mask), N.ptr(self.msum), N.ptr(self.status),
            N.ptr(self.overlap), N.ptr(self.gsum), N.ptr(self.f), N.ptr(self.avg_f), N.ptr(self.max_f),
            N.ptr(self.ws), self.ws_bytes, N.current_stream()))
        return self

    def knapsack(self, values):
        """Stand-alone knapsack over explicit int32 values (utils/knapsack.py:5-23)."""
        values = self._to_dev(values, torch.int32)
        N.check(N.lib().smz_knapsack(
            N.ptr(self.d_desc), self.n_videos, N.ptr(values), N.ptr(self.d_nfps), self.max_n_segs,
            self.max_capacity, self.max_n_frames, self.max_seg_frames, N.ptr(self.picked), N.ptr(self.mask), N.ptr(self.msum),
            N.ptr(self.status), N.ptr(self.ws), self.ws_bytes, N.current_stream()))
        return self

    def pack_summary(self, machine):
        """Use an explicit machine summary (packed like self.summary) for the next fscore()."""
        machine = self._to_dev(machine, torch.float32)
        N.check(N.lib().smz_pack_summary(N.ptr(self.d_desc), self.n_videos, self.max_n_frames, N.ptr(machine),
                                         N.ptr(self.mask), N.ptr(self.msum), N.current_stream()))
        return self

    # ---- annotator summaries staged as 1 bit per frame (host -> device copies move 32x fewer bytes) -------------
    def _bits_layout(self):
        if not hasattr(self, "h_bits_off"):
            words = ((self.h_desc["n_frames"].astype(np.int64) + 31) // 32) * self.h_desc["n_users"].astype(np.int64)
            off = np.zeros(self.n_videos, dtype=np.int64)
            if self.n_videos:
                off[1:] = np.cumsum(words)[:-1]
            self.h_bits_off, self.total_bit_words = off, int(words.sum())
            self.d_bits_off = torch.from_numpy(off).to(self.device)
        return self.h_bits_off

    def pack_user_summary_host(self, h_users, out=None, n_threads=None):
        """HOST: (x > 0) of the float32 annotator rows (flat, laid out like ``d_users``) -> int32 words, one bit per
        frame (``evaluate_summary`` binarises exactly so, utils/eval.py:148-149).  ``out``: optional (pinned) int32
        tensor of ``total_bit_words``; returns it.  Runs on host threads inside the caller's thread (the GIL is
        released), typically one step ahead of the copy."""
        self._bits_layout()
        if h_users.dtype != torch.float32 or h_users.is_cuda or not h_users.is_contiguous():
            raise ValueError("pack_user_summary_host: contiguous float32 host tensor expected")
        if out is None:
            out = torch.empty(max(self.total_bit_words, 1), dtype=torch.int32)
        if n_threads is None:
            import os
            n_threads = min(os.cpu_count() or 1, 32)
        N.check(N.lib().smz_host_pack_user_summary(self.h_desc.ctypes.data_as(ctypes.c_void_p), self.n_videos,
                                                   ctypes.c_void_p(h_users.data_ptr()),
                                                   self.h_bits_off.ctypes.data_as(ctypes.c_void_p),
                                                   ctypes.c_void_p(out.data_ptr()), int(n_threads)))
        return out

    def pack_user_bits(self, out=None):
        """DEVICE: the resident float32 annotator rows -> int32 words, one bit per frame (same layout and bits as
        ``pack_user_summary_host``).  HBM-bound streaming kernel (smz_pack_user_bits)."""
        if not self.has_users:
            raise ValueError("this batch holds no user_summary")
        self._bits_layout()
        if out is None:
            out = torch.empty(max(self.total_bit_words, 1), dtype=torch.int32, device=self.device)
        N.check(N.lib().smz_pack_user_bits(N.ptr(self.d_desc), self.n_videos, N.ptr(self.d_users), N.ptr(self.d_bits_off),
                                           N.ptr(out), N.current_stream()))
        return out

    def fscore_packed(self, d_bits):
        """fscore() against annotator rows given as device bit words (see pack_user_summary_host); same F values."""
        self._bits_layout()
        if d_bits.numel() < self.total_bit_words or d_bits.dtype != torch.int32 or not d_bits.is_cuda:
            raise ValueError("fscore_packed: int32 device tensor of total_bit_words expected")
        N.check(N.lib().smz_fscore_packed(
            N.ptr(self.d_desc), self.n_videos, N.ptr(d_bits), N.ptr(self.d_bits_off), N.ptr(self.mask), N.ptr(self.msum),
            N.ptr(self.overlap), N.ptr(self.gsum), N.ptr(self.f), N.ptr(self.avg_f), N.ptr(self.max_f), N.current_stream()))
        return self

    def fscore(self):
        """evaluate_summary for every video (utils/eval.py:125-165) against the resident
        user summaries, using the mask of the last select()/pack_summary()."""
        if not self.has_users:
            raise ValueError("this batch holds no user_summary")
        N.check(N.lib().smz_fscore(
            N.ptr(self.d_desc), self.n_videos, self.max_n_frames, self.total_users, N.ptr(self.d_users),
            N.ptr(self.mask), N.ptr(self.msum), N.ptr(self.overlap), N.ptr(self.gsum), N.ptr(self.f),
            N.ptr(self.avg_f), N.ptr(self.max_f), N.current_stream()))
        return self

    def upsample(self, scores):
        """upsample for every video (utils/eval.py:15-35) -> packed (sum n_frames,) device tensor."""
        scores = self._to_dev(scores, torch.float32)
        out = torch.empty(max(self.total_frames, 1), dtype=torch.float32, device=self.device)
        N.check(N.lib().smz_upsample(N.ptr(self.d_desc), self.n_videos, self.max_n_frames, N.ptr(scores),
                                     N.ptr(self.d_picks), N.ptr(out), N.ptr(self.status), N.current_stream()))
        return out

    def check_status(self):
        st = self.status[: self.n_videos].cpu().numpy()
        bad = np.nonzero(st)[0]
        if bad.size:
            v = int(bad[0])
            if st[v] & N.SMZ_STATUS_INTERVALS:
                raise IndexError(f"video {v}: more upsample intervals than scores (utils/eval.py:29-34)")
            if st[v] & N.SMZ_STATUS_WEIGHT_RANGE:
                raise ValueError(f"video {v}: a segment is longer than the declared max_seg_frames")
            raise OverflowError(f"video {v}: segment values exceed the int32 knapsack range")

    # host views -----------------------------------------------------------------------------
    def summary_of(self, i):
        d = self.h_desc[i]
        return self.summary[int(d["summ_off"]): int(d["summ_off"]) + int(d["summ_len"])]

    def picked_of(self, i):
        d = self.h_desc[i]
        return self.picked[int(d["seg_off"]): int(d["seg_off"]) + int(d["n_segs"])]

    def users_slice(self, i):
        d = self.h_desc[i]
        return slice(int(d["ucount_off"]), int(d["ucount_off"]) + int(d["n_users"]))

"""Deterministic synthetic SumMe-/TVSum-/sweep-shaped datasets.

The real ``summarizer_dataset_*_google_pool5.h5`` files are not distributable (and there is no
network), so tests and benchmarks run on seeded synthetic videos that follow the HDF5 schema of
the reference (datasets/README.md:5-42) field by field.  ``ArrayDataset`` exposes the h5py access
patterns the reference Trainer uses (``ds[key]["features"][...]``, ``d["n_frames"][()]``,
``"user_scores" in d``, ``ds.keys()``).
"""
import numpy as np

DATASET_ID = {"summe": 1, "tvsum": 2, "sweep": 5}


class _Field:
    __slots__ = ("value",)

    def __init__(self, value):
        self.value = value

    def __getitem__(self, idx):
        if idx is Ellipsis:
            return np.array(self.value, copy=True)
        if isinstance(idx, tuple) and len(idx) == 0:
            return self.value
        return np.asarray(self.value)[idx]

    @property
    def shape(self):
        return np.shape(self.value)


class _Group(dict):
    def __getitem__(self, k):
        return _Field(dict.__getitem__(self, k))

    def raw(self, k):
        return dict.__getitem__(self, k)


class ArrayDataset:
    """In-memory stand-in for the read side of ``h5py.File`` (models/__init__.py:15)."""

    def __init__(self, videos, name="synthetic"):
        self._videos = {k: _Group(v) for k, v in videos.items()}
        self.name = name

    def __getitem__(self, key):
        return self._videos[key]

    def __contains__(self, key):
        return key in self._videos

    def keys(self):
        return self._videos.keys()

    def __len__(self):
        return len(self._videos)

    def close(self):
        pass


def _segments(rng, n_frames, uniform=None):
    """Change points: inclusive [start, end] rows; lengths i.i.d. uniform in [30, 300] frames, or a
    fixed ``uniform`` length (the paper's 2-second segmentation for Twitch-LOL), last one truncated."""
    cps, s = [], 0
    while s < n_frames:
        length = int(uniform) if uniform else int(rng.integers(30, 301))
        e = min(s + length - 1, n_frames - 1)
        cps.append((s, e))
        s = e + 1
    cps = np.asarray(cps, dtype=np.int32)
    return cps, (cps[:, 1] - cps[:, 0] + 1).astype(np.int32)


def _user_summary(rng, nfps, n_frames, n_users, proportion=0.15):
    """Shot-aligned 0/1 annotations: every annotator takes segments in a private random order
    while they fit a ``proportion`` budget (≈15 % ones, like knapsack-made ground truth)."""
    budget = int(n_frames * proportion)
    out = np.zeros((n_users, n_frames), dtype=np.float32)
    starts = np.concatenate([[0], np.cumsum(nfps)[:-1]])
    for u in range(n_users):
        total = 0
        for s in rng.permutation(len(nfps)):
            if total + nfps[s] <= budget:
                out[u, starts[s]:starts[s] + nfps[s]] = 1.0
                total += int(nfps[s])
    return out


def make_video(dataset, index, n_frames=None, n_users=None, uniform_segments=None, feature_dim=1024,
               with_features=True):
    """One synthetic video (all fields of datasets/README.md:5-42); seed = 1000*dataset_id + index."""
    did = DATASET_ID[dataset]
    rng = np.random.default_rng(1000 * did + index)
    if n_frames is None:
        if dataset == "summe":
            n_frames = 4494 if index == 1 else int(rng.integers(950, 9722))
        elif dataset == "tvsum":
            n_frames = 10597 if index == 1 else int(rng.integers(2500, 19407))
        else:
            n_frames = 30000
    if n_users is None:
        n_users = 15 + (index % 4) if dataset == "summe" else 20
    picks = np.arange(0, n_frames, 15, dtype=np.int32)
    n_steps = len(picks)
    v = {"n_frames": np.int32(n_frames), "n_steps": np.int32(n_steps), "picks": picks,
         "video_name": f"{dataset}_synthetic_{index}"}
    if with_features:
        feat = np.abs(rng.standard_normal((n_steps, feature_dim), dtype=np.float32))
        feat /= np.linalg.norm(feat, axis=1, keepdims=True)
        v["features"] = feat.astype(np.float32)
    raw = rng.random(n_steps + 8, dtype=np.float32)
    gt = np.convolve(raw, np.ones(9, dtype=np.float32) / 9.0, mode="valid")[:n_steps].astype(np.float32)
    if gt.max() == gt.min():
        gt[0] += 0.5
    v["gtscore"] = gt
    cps, nfps = _segments(rng, n_frames, uniform_segments)
    v["change_points"], v["n_frame_per_seg"] = cps, nfps
    v["user_summary"] = _user_summary(rng, nfps, n_frames, n_users)
    v["gtsummary"] = (gt > np.quantile(gt, 0.85)).astype(np.float32)
    if dataset == "summe":
        v["user_scores"] = np.repeat(gt, 15)[:n_frames][None, :].astype(np.float32)   # normalize_datasets.py:53-59
    else:
        blocks = (n_frames + 59) // 60
        lv = rng.integers(0, 5, size=(n_users, blocks)).astype(np.float32) / 4.0        # (anno-1)/4, :25
        v["user_scores"] = np.repeat(lv, 60, axis=1)[:, :n_frames]
    return v


def make_dataset(dataset, n_videos=None, **kw):
    """SumMe-shaped (25 videos, keys video_1..25) or TVSum-shaped (50 videos, video_1..50)."""
    if n_videos is None:
        n_videos = {"summe": 25, "tvsum": 50}[dataset]
    return ArrayDataset({f"video_{i}": make_video(dataset, i, **kw) for i in range(1, n_videos + 1)}, name=dataset)


def write_dataset_h5(dataset, path):
    """Writes an ``ArrayDataset`` as an HDF5 file with the schema of datasets/README.md:5-42 (one group per video key,
    scalars as 0-d datasets, ``video_name`` as a string) — with h5py when it is installed, else with utils/hdf5.py."""
    try:
        import h5py
        opener = h5py.File
    except ImportError:
        from .utils import hdf5
        opener = hdf5.File
    with opener(path, "w") as f:
        for key in dataset.keys():
            g = f.create_group(key)
            for name in dict.keys(dataset[key]):
                val = dataset[key].raw(name)
                g.create_dataset(name, data=np.bytes_(val) if isinstance(val, str) else val)
            if "video_name" not in dataset[key]:
                g.create_dataset("video_name", data=np.bytes_(f"{dataset.name}_{key}"))
    return path


# ----------------------------------------------------------------------------------------------
# sweep (config 5): 10k videos x 2000 steps (30 000 frames), 20 annotators — generated ON the device
# ----------------------------------------------------------------------------------------------
def make_sweep_batch(n_videos, device, n_frames=30000, n_users=20, proportion=0.15, seed=5000,
                     chunk=128):
    """Builds a resident ``VideoBatch`` for the sweep without staging 2.4 MB/video through the host.
    Segment structure comes from numpy (seeded per video); annotator summaries are produced on the
    GPU with torch ops (random-priority greedy selection under the 15 % budget)."""
    import torch
    from . import _native as N
    from .batch import VideoBatch, capacity_of

    ld = (n_frames + 3) // 4 * 4
    desc = np.zeros(n_videos, dtype=N.VIDEO_DESC)
    picks1 = np.arange(0, n_frames, 15, dtype=np.int32)
    n_steps = len(picks1)
    cps_all, nfps_all = [], []
    sg = 0
    for i in range(n_videos):
        rng = np.random.default_rng(seed + i)
        cps, nfps = _segments(rng, n_frames, 60 if i % 16 == 15 else None)
        d = desc[i]
        d["score_off"], d["picks_off"], d["seg_off"] = i * n_steps, i * n_steps, sg
        d["user_off"], d["user_ld"] = i * n_users * ld, ld
        d["summ_off"], d["frame_off"] = i * n_frames, i * n_frames
        d["mask_off"], d["ucount_off"] = i * ((n_frames + 31) // 32), i * n_users
        d["n_scores"] = d["n_picks"] = n_steps
        d["n_segs"], d["n_users"], d["n_frames"], d["summ_len"] = len(nfps), n_users, n_frames, n_frames
        d["capacity"] = capacity_of(n_frames, proportion)
        cps_all.append(cps.reshape(-1)); nfps_all.append(nfps)
        sg += len(nfps)
    max_segs = max(len(x) for x in nfps_all)
    seg_len = np.zeros((n_videos, max_segs), dtype=np.int64)
    for i, w in enumerate(nfps_all):
        seg_len[i, :len(w)] = w
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    users = torch.zeros(n_videos * n_users * ld, dtype=torch.float32, device=dev)
    budget = int(n_frames * proportion)
    for v0 in range(0, n_videos, chunk):
        v1 = min(v0 + chunk, n_videos)
        L = torch.from_numpy(seg_len[v0:v1]).to(dev)                                  # [V,S]
        V = v1 - v0
        pri = torch.rand((V, n_users, max_segs), generator=g, device=dev)
        pri = pri.masked_fill((L == 0)[:, None, :], -1.0)
        order = pri.argsort(dim=2, descending=True)
        Ls = torch.gather(L[:, None, :].expand(V, n_users, max_segs), 2, order)
        take_sorted = (Ls.cumsum(dim=2) <= budget) & (Ls > 0)
        take = torch.zeros_like(take_sorted)
        take.scatter_(2, order, take_sorted)
        flags = take.to(torch.float32).reshape(-1)
        reps = L[:, None, :].expand(V, n_users, max_segs).reshape(-1)
        frames = torch.repeat_interleave(flags, reps).reshape(V * n_users, n_frames)
        users[v0 * n_users * ld: v1 * n_users * ld].view(V * n_users, ld)[:, :n_frames] = frames
    picks = np.tile(picks1, n_videos)
    return VideoBatch.from_packed(desc, picks, np.concatenate(cps_all), np.concatenate(nfps_all), users,
                                  proportion=proportion, device=dev)

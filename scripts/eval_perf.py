"""Evaluation-path timing on sweep-shaped videos: selection alone, F-score alone, and the pipelined smz_eval_batch
for several slice counts (CUDA events, L2-exceeding inputs).  python scripts/eval_perf.py [n_videos]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from summarizer_b200 import synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
dev = torch.device("cuda", 0)
b = synthetic.make_sweep_batch(n, dev, seed=5000)
g = torch.Generator(device=dev); g.manual_seed(1)
mode = os.environ.get("SCORES", "uniform")
if mode == "uniform":
    scores = torch.rand(b.total_scores, generator=g, device=dev)
else:                                  # VASNet-like: sigmoid outputs in a narrow band
    scores = torch.sigmoid(0.3 * torch.randn(b.total_scores, generator=g, device=dev) + 0.2)
d = b.h_desc
b_eval = int((4 * d["n_users"].astype(np.int64) * d["n_frames"] + 8 * d["n_scores"] + 12 * d["n_segs"] + 4 * d["n_frames"]
              + 12 * d["n_users"]).sum())


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    e[0].record()
    for i in range(reps):
        fn(); e[i + 1].record()
    torch.cuda.synchronize()
    return float(np.median([e[i].elapsed_time(e[i + 1]) for i in range(reps)]))


out = {"videos": n, "scores": mode, "no_dp16": bool(os.environ.get("SMZ_NO_DP16"))}
out["select_ms"] = timed(lambda: b.select(scores))
out["fscore_ms"] = timed(lambda: b.fscore())
pk = b.pack_user_bits()
out["pack_ms"] = timed(lambda: b.pack_user_bits(pk))
ref = {k: getattr(b, k).clone() for k in ("picked", "mask", "msum", "overlap", "avg_f")}
out["evaluate_ms"] = timed(lambda: b.evaluate(scores))
torch.cuda.synchronize()
assert all(torch.equal(ref[k], getattr(b, k)) for k in ref)
best = out["evaluate_ms"]
out["eval_path_frac_of_6541.8GBs"] = b_eval / (best / 1e3) / 1e9 / 6541.8
fb = b.ws.view(torch.int32)
b.check_status()
print(json.dumps(out))

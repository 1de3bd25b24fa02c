"""GPU parity tests of the shot-selection / F-score kernels (through the C ABI) against the oracle
and the golden vectors produced by the unmodified reference.  Integer results must be bit-exact;
float32 results the reference computes in float32 must be bit-exact too."""
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle, eval_np as E

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "eval_golden.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def _same(a, b):
    return np.float64(a) == np.float64(b)


# ------------------------------------------------------------------------------------------------
# per-video API (reference signatures) against golden vectors from the reference itself
# ------------------------------------------------------------------------------------------------
def test_golden_per_video_api(golden):
    from summarizer_b200.utils import eval as P
    for name in golden["names"]:
        g = lambda k: golden[f"{name}/{k}"]
        nf = int(g("n_frames"))
        us = g("user_summary").astype(np.float32)
        fs = P.upsample(g("scores"), nf, g("picks"))
        assert fs.dtype == np.float32 and np.array_equal(fs, g("frame_scores")), name
        seg = E.segment_scores(g("frame_scores"), g("cps"))
        for method in ("knapsack", "rank"):
            if method == "rank" and len(np.unique(seg)) < len(seg):
                continue  # ties decided by numpy's unstable argsort in the reference: unpinned
            s = P.generate_summary(g("scores"), g("cps"), nf, g("nfps").tolist(), g("picks"), 0.15, method)
            assert s.dtype == np.float32
            assert np.array_equal(s.astype(np.uint8), g(f"summary_{method}")), (name, method)
            avg_f, max_f = P.evaluate_summary(s, us)
            assert _same(avg_f, g(f"f_{method}")[0]) and _same(max_f, g(f"f_{method}")[1]), (name, method)
        s = g("summary_knapsack").astype(np.float32)
        for tag, m in (("long", np.concatenate([s, np.ones(7, np.float32)])), ("short", s[: nf - 11])):
            avg_f, max_f = P.evaluate_summary(m, us)
            assert _same(avg_f, g(f"f_{tag}")[0]) and _same(max_f, g(f"f_{tag}")[1]), (name, tag)


def test_unknown_method_and_metric_raise():
    from summarizer_b200.utils import eval as P
    with pytest.raises(KeyError):
        P.generate_summary(np.ones(4, np.float32), np.array([[0, 59]]), 60, [60], np.arange(0, 60, 15), 0.15, "nope")
    with pytest.raises(KeyError):
        P.evaluate_scores(np.ones(4), np.ones((1, 4)), metric="nope")


# ------------------------------------------------------------------------------------------------
# batched path against the C oracle
# ------------------------------------------------------------------------------------------------
def _check_batch_against_oracle(videos, scores_list, method="knapsack", pad_user_rows=True, sample=None):
    from summarizer_b200.batch import VideoBatch
    b = VideoBatch(videos, proportion=0.15, pad_user_rows=pad_user_rows)
    scores = torch.from_numpy(np.concatenate(scores_list).astype(np.float32))
    b.select(scores, method=method)
    b.check_status()
    has_users = b.has_users
    if has_users:
        b.fscore()
    torch.cuda.synchronize()
    seg_mean = b.seg_mean.cpu().numpy(); values = b.values.cpu().numpy(); picked = b.picked.cpu().numpy()
    summary = b.summary.cpu().numpy(); mask = b.mask.cpu().numpy().view(np.uint32); msum = b.msum.cpu().numpy()
    overlap = b.overlap.cpu().numpy(); gsum = b.gsum.cpu().numpy(); f = b.f.cpu().numpy()
    avg_f = b.avg_f.cpu().numpy(); max_f = b.max_f.cpu().numpy()
    idx = range(len(videos)) if sample is None else sample
    for i in idx:
        v, d = videos[i], b.h_desc[i]
        ref_sum, parts = c_oracle.generate_summary(scores_list[i], v["change_points"], int(v["n_frames"]),
                                                   v["n_frame_per_seg"], v["picks"], 0.15, method)
        so, ns = int(d["seg_off"]), int(d["n_segs"])
        assert seg_mean[so:so + ns].tobytes() == parts["seg_score"].tobytes(), i
        assert np.array_equal(values[so:so + ns], parts["values"]), i
        assert np.nonzero(picked[so:so + ns])[0].tolist() == parts["picks"], i
        assert int(d["capacity"]) == parts["capacity"]
        mo, sl = int(d["summ_off"]), int(d["summ_len"])
        assert np.array_equal(summary[mo:mo + sl], ref_sum), i
        nf = int(d["n_frames"])
        bits = np.unpackbits(mask[int(d["mask_off"]):int(d["mask_off"]) + (nf + 31) // 32].view(np.uint8),
                             bitorder="little")[:nf]
        m_ref = np.zeros(nf, np.uint8); lim = min(nf, sl); m_ref[:lim] = ref_sum[:lim] > 0
        assert np.array_equal(bits, m_ref), i
        assert msum[i] == int(m_ref.sum())
        if has_users:
            r = c_oracle.evaluate_summary(ref_sum, v["user_summary"])
            uo, nu = int(d["ucount_off"]), int(d["n_users"])
            assert np.array_equal(overlap[uo:uo + nu], r["overlap"]), i
            assert np.array_equal(gsum[uo:uo + nu], r["gsum"]), i
            assert f[uo:uo + nu].tobytes() == r["f"].tobytes(), i
            assert avg_f[i] == r["avg_f"] and max_f[i] == r["max_f"], i
            # and the numpy restatement (which follows numpy's dtype promotion, including the float64 case of a
            # summary shorter than n_frames, utils/eval.py:143-145) agrees
            a2, m2 = E.evaluate_summary(ref_sum, v["user_summary"])
            assert _same(avg_f[i], a2) and _same(max_f[i], m2), i
    return b


def _dataset_videos(name, n, **kw):
    from summarizer_b200 import synthetic
    ds = synthetic.make_dataset(name, n, with_features=False, **kw)
    vids = [{k: ds[key].raw(k) for k in ("n_frames", "picks", "change_points", "n_frame_per_seg", "user_summary")}
            for key in ds.keys()]
    rng = np.random.default_rng(7)
    scores = [rng.random(len(v["picks"])).astype(np.float32) for v in vids]
    return vids, scores


def test_summe_shaped_batch_knapsack():
    vids, scores = _dataset_videos("summe", 25)
    _check_batch_against_oracle(vids, scores, "knapsack")


def test_tvsum_shaped_batch_knapsack_and_rank():
    vids, scores = _dataset_videos("tvsum", 50)          # the whole TVSum-shaped dataset
    _check_batch_against_oracle(vids, scores, "knapsack")
    _check_batch_against_oracle(vids, scores, "rank")


def test_uniform_segments_tie_heavy():
    vids, _ = _dataset_videos("tvsum", 10, uniform_segments=60)
    rng = np.random.default_rng(8)
    scores = [(rng.integers(0, 3, len(v["picks"])) / 2.0).astype(np.float32) for v in vids]   # values in {0,500,1000}
    _check_batch_against_oracle(vids, scores, "knapsack")


def test_unaligned_user_rows_scalar_path():
    vids, scores = _dataset_videos("summe", 6)
    for v in vids:   # make n_frames odd so unpadded rows are misaligned
        assert v["user_summary"].shape[1] == int(v["n_frames"])
    _check_batch_against_oracle(vids, scores, "knapsack", pad_user_rows=False)


def test_edge_cases():
    from summarizer_b200 import synthetic
    vids, scores = [], []
    rng = np.random.default_rng(9)
    # tiny video: capacity 0; single segment; one annotator; all-fit shortcut (proportion irrelevant here)
    for nf, uni, nu in [(5, 5, 1), (61, 61, 2), (100, 10, 1), (2049, 33, 3), (4097, 2048, 2), (33, 1, 1)]:
        v = synthetic.make_video("summe", 300 + nf, n_frames=nf, n_users=nu, uniform_segments=uni, with_features=False)
        vids.append({k: v[k] for k in ("n_frames", "picks", "change_points", "n_frame_per_seg", "user_summary")})
        scores.append(rng.random(len(v["picks"])).astype(np.float32))
    # all-zero scores: every value 0 -> the OR-tools item-0 default-id quirk decides
    v = synthetic.make_video("summe", 400, n_frames=3000, n_users=2, with_features=False)
    vids.append({k: v[k] for k in ("n_frames", "picks", "change_points", "n_frame_per_seg", "user_summary")})
    scores.append(np.zeros(len(v["picks"]), np.float32))
    # annotator with no selected frame at all (F = 0 -> float64 promotion in the reference)
    v = synthetic.make_video("summe", 401, n_frames=2500, n_users=3, with_features=False)
    v["user_summary"][1] = 0
    vids.append({k: v[k] for k in ("n_frames", "picks", "change_points", "n_frame_per_seg", "user_summary")})
    scores.append(rng.random(len(v["picks"])).astype(np.float32))
    _check_batch_against_oracle(vids, scores, "knapsack")
    _check_batch_against_oracle(vids, scores, "rank")


def test_summary_longer_and_shorter_than_n_frames():
    """sum(nfps) != n_frames: the summary vector keeps sum(nfps) entries, the mask is truncated/padded."""
    from summarizer_b200 import synthetic
    vids, scores = [], []
    rng = np.random.default_rng(10)
    for k, delta in enumerate((+13, -9)):
        v = synthetic.make_video("summe", 500 + k, n_frames=3000, n_users=2, with_features=False)
        nfps = v["n_frame_per_seg"].copy(); nfps[-1] += delta
        vids.append(dict(n_frames=v["n_frames"], picks=v["picks"], change_points=v["change_points"],
                         n_frame_per_seg=nfps, user_summary=v["user_summary"]))
        scores.append(rng.random(len(v["picks"])).astype(np.float32))
    _check_batch_against_oracle(vids, scores, "knapsack")


def test_large_video_bits_in_global_workspace():
    """LOL-sized video: the take-bit matrix does not fit shared memory -> work-buffer path."""
    from summarizer_b200 import synthetic
    from summarizer_b200.batch import VideoBatch
    v = synthetic.make_video("tvsum", 600, n_frames=100000, n_users=2, uniform_segments=60, with_features=False)
    rng = np.random.default_rng(11)
    sc = rng.random(len(v["picks"])).astype(np.float32)
    vid = {k: v[k] for k in ("n_frames", "picks", "change_points", "n_frame_per_seg", "user_summary")}
    b = VideoBatch([vid, vid])
    assert b.ws_bytes > 0
    b.select(torch.from_numpy(np.concatenate([sc, sc])))
    b.check_status()
    fs = E.upsample(sc, 100000, v["picks"])
    seg = E.segment_scores(fs, v["change_points"])
    ref = E.knapsack_dp_takebits(E.knapsack_values(seg), v["n_frame_per_seg"], E.capacity_of(100000, 0.15))
    n = len(seg)
    picked = b.picked.cpu().numpy()
    assert np.nonzero(picked[:n])[0].tolist() == ref
    assert np.nonzero(picked[n:2 * n])[0].tolist() == ref


def test_long_video_register_dp_with_cursor_pooling():
    """70 000 frames: the DP row still fits the register kernel (K = 54) but the frame staging of the pooling kernel
    does not (>= 65 535 frames) -> the cursor-walking pool_kernel; the second video has more than 128*8 frames per
    segment on average in its tail, so the out-of-line deep pairwise recursion runs as well."""
    from summarizer_b200 import synthetic
    from summarizer_b200.batch import VideoBatch
    v = synthetic.make_video("tvsum", 601, n_frames=70000, n_users=2, with_features=False)
    w = synthetic.make_video("tvsum", 602, n_frames=70000, n_users=2, uniform_segments=2500, with_features=False)
    rng = np.random.default_rng(12)
    vids, scs = [], []
    for x in (v, w):
        vids.append({k: x[k] for k in ("n_frames", "picks", "change_points", "n_frame_per_seg", "user_summary")})
        scs.append(rng.random(len(x["picks"])).astype(np.float32))
    b = VideoBatch(vids)
    b.select(torch.from_numpy(np.concatenate(scs)))
    b.check_status()
    picked, off = b.picked.cpu().numpy(), 0
    for x, sc in zip((v, w), scs):
        seg = E.segment_scores(E.upsample(sc, 70000, x["picks"]), x["change_points"])
        ref = E.knapsack_dp_takebits(E.knapsack_values(seg), x["n_frame_per_seg"], E.capacity_of(70000, 0.15))
        n = len(seg)
        assert np.nonzero(picked[off:off + n])[0].tolist() == ref
        off += n


def test_fscore_from_host_packed_annotator_bits():
    """smz_host_pack_user_summary + smz_fscore_packed: the 1-bit-per-frame staging form gives bit-identical F values
    (evaluate_summary binarises with x > 0 first, utils/eval.py:148-149)."""
    from summarizer_b200.batch import VideoBatch
    vids, scores = _dataset_videos("tvsum", 6)
    vids[2] = dict(vids[2]); vids[2]["user_summary"] = vids[2]["user_summary"] * np.float32(0.25)   # non-binary positives
    vids[3] = dict(vids[3]); vids[3]["user_summary"] = vids[3]["user_summary"] - np.float32(0.5)    # negatives count as 0
    b = VideoBatch(vids)
    b.select(torch.from_numpy(np.concatenate(scores)))
    b.fscore()
    ref = [t.clone() for t in (b.overlap, b.gsum, b.f, b.avg_f, b.max_f)]
    h_users = b.d_users.cpu()
    h_bits = b.pack_user_summary_host(h_users, n_threads=3)
    assert h_bits.numel() == b.total_bit_words
    for t in (b.overlap, b.gsum, b.f, b.avg_f, b.max_f):
        t.zero_()
    b.fscore_packed(h_bits.cuda())
    for r, t in zip(ref, (b.overlap, b.gsum, b.f, b.avg_f, b.max_f)):
        assert torch.equal(r[: b.total_users if r.numel() >= b.total_users else b.n_videos], t[: b.total_users if r.numel() >= b.total_users else b.n_videos])


@pytest.mark.gpu
def test_device_pack_equals_host_pack():
    """smz_pack_user_bits (bulk-copy streaming kernel) produces the words smz_host_pack_user_summary does: aligned and
    unaligned rows, ragged lengths (chunk tails), non-binary / negative annotations, more chunks than ring stages."""
    from summarizer_b200 import synthetic
    from summarizer_b200.batch import VideoBatch
    vids, _ = _dataset_videos("tvsum", 5)
    vids[1] = dict(vids[1]); vids[1]["user_summary"] = vids[1]["user_summary"] * np.float32(0.25)
    vids[2] = dict(vids[2]); vids[2]["user_summary"] = vids[2]["user_summary"] - np.float32(0.5)
    for nf, nu in ((5, 1), (1024, 2), (1025, 3), (33000, 7), (4097, 21)):
        v = synthetic.make_video("summe", 700 + nf, n_frames=nf, n_users=nu, with_features=False)
        vids.append({k: v[k] for k in ("n_frames", "picks", "change_points", "n_frame_per_seg", "user_summary")})
    for pad in (True, False):
        b = VideoBatch(vids, pad_user_rows=pad)
        want = b.pack_user_summary_host(b.d_users.cpu(), n_threads=2)
        got = b.pack_user_bits(torch.full((b.total_bit_words,), -1, dtype=torch.int32, device="cuda"))
        assert torch.equal(got.cpu(), want), f"pad_user_rows={pad}"


def test_knapsack_ortools_signature():
    from summarizer_b200.utils.knapsack import knapsack_ortools
    rng = np.random.default_rng(12)
    for _ in range(40):
        n = int(rng.integers(1, 40))
        w = rng.integers(1, 60, n)
        vals = rng.integers(0, 5, n) / 4.0 if rng.random() < 0.5 else rng.random(n)
        cap = int(rng.integers(0, 400))
        got = knapsack_ortools(vals.tolist(), w.tolist(), n, cap)
        assert got == E.knapsack_ortools(vals, w, n, cap)
    assert knapsack_ortools([0.0, 0.0, 0.5], [3, 4, 5], 3, 12) == [0, 1, 2]      # sum(w) <= capacity
    assert knapsack_ortools([], [], 0, 10) == []


def test_sweep_shaped_invariants_and_sample():
    """Config-5-shaped videos (30 000 frames, 20 annotators) generated on the device: size-independent
    properties on all videos plus exact comparison of a sample with the C oracle."""
    from summarizer_b200 import synthetic
    n = 96
    b = synthetic.make_sweep_batch(n, "cuda", seed=5000)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    scores = torch.rand(b.total_scores, generator=g, device="cuda")
    b.select(scores); b.fscore(); b.check_status()
    torch.cuda.synchronize()
    msum = b.msum.cpu().numpy(); picked = b.picked.cpu().numpy(); nfps = b.d_nfps.cpu().numpy()
    overlap = b.overlap.cpu().numpy().reshape(n, 20); gsum = b.gsum.cpu().numpy().reshape(n, 20)
    users = b.d_users.view(n, 20, -1)
    assert np.array_equal(gsum, (users > 0).sum(dim=2).cpu().numpy())
    summ = b.summary.view(n, 30000)
    assert np.array_equal(msum, summ.sum(dim=1).cpu().numpy().astype(np.int64))
    assert np.array_equal(overlap, ((users[:, :, :30000] > 0) & (summ[:, None, :] > 0)).sum(dim=2).cpu().numpy())
    for i in range(n):
        d = b.h_desc[i]
        so, ns = int(d["seg_off"]), int(d["n_segs"])
        sel_len = int(nfps[so:so + ns][picked[so:so + ns] > 0].sum())
        assert sel_len == msum[i] and sel_len <= int(d["capacity"])
    # exact sample
    sc = scores.cpu().numpy(); cps = b.d_cps.cpu().numpy().reshape(-1, 2); picks = b.d_picks.cpu().numpy()
    for i in (0, 15, 47, 95):
        d = b.h_desc[i]
        so, ns = int(d["seg_off"]), int(d["n_segs"])
        ref_sum, parts = c_oracle.generate_summary(sc[int(d["score_off"]):int(d["score_off"]) + 2000], cps[so:so + ns],
                                                   30000, nfps[so:so + ns], picks[int(d["picks_off"]):int(d["picks_off"]) + 2000])
        assert np.nonzero(picked[so:so + ns])[0].tolist() == parts["picks"]
        assert np.array_equal(summ[i].cpu().numpy(), ref_sum)


# ------------------------------------------------------------------------------------------------
# 16-bit knapsack DP (dp16_kernel) and its hand-over to the 32-bit kernel; sliced two-stream evaluation
# ------------------------------------------------------------------------------------------------
def _knapsack_batch(cases):
    """cases: list of (values int array, weights int array, capacity) -> picked index lists from smz_knapsack."""
    from summarizer_b200 import _native as N
    from summarizer_b200.batch import VideoBatch
    desc = np.zeros(len(cases), dtype=N.VIDEO_DESC)
    so = 0
    for i, (vals, w, cap) in enumerate(cases):
        desc[i]["seg_off"], desc[i]["n_segs"], desc[i]["capacity"] = so, len(w), cap
        so += len(w)
    nfps = np.concatenate([np.asarray(c[1], np.int32) for c in cases])
    b = VideoBatch.from_packed(desc, np.zeros(1, np.int32), np.zeros(2, np.int32), nfps, None, 0.15)
    b.knapsack(torch.from_numpy(np.concatenate([np.asarray(c[0], np.int32) for c in cases])))
    torch.cuda.synchronize()
    picked, st, out, so = b.picked.cpu().numpy(), b.status.cpu().numpy(), [], 0
    for vals, w, cap in cases:
        out.append(np.nonzero(picked[so:so + len(w)])[0].tolist()); so += len(w)
    return out, st


def test_dp16_matches_oracle_on_every_hand_over_condition():
    """Random instances built to hit each branch of dp16_kernel: plain 16-bit runs, values above 32767, negative values,
    row values that outgrow 16 bits half way (abort -> 32-bit redo), weights above the mirror pad, zero weights, zero
    values (item-0 quirk), everything-fits and nothing-fits shortcuts.  Bit-exact against the C oracle."""
    rng = np.random.default_rng(21)
    cases = []
    for kind in range(12):
        for rep in range(6):
            n = int(rng.integers(1, 200))
            w = rng.integers(1, 300, n)
            vals = rng.integers(0, 1001, n)
            cap = int(rng.integers(1, 4500))
            if kind == 1: vals = rng.integers(0, 40000, n)                   # p > 32767 somewhere: 32-bit path
            if kind == 2: vals = rng.integers(-500, 1000, n)                 # negative values: 32-bit path
            if kind == 3: vals = rng.integers(2000, 6000, n); cap = 4400     # overflows 16 bits while running: abort
            if kind == 4: w = rng.integers(1, 2500, n)                       # items heavier than 512: masked variant
            if kind == 5: w = rng.integers(0, 3, n); cap = int(rng.integers(1, 40))   # zero weights, tiny capacity
            if kind == 6: vals = np.zeros(n, np.int64)                       # all-zero values: default id 0 quirk
            if kind == 7: cap = int(w.sum()) + 3                             # everything fits
            if kind == 8: w = rng.integers(5000, 6000, n)                    # nothing fits
            if kind == 9: vals = rng.integers(0, 3, n) * 500; w = np.full(n, 60)   # tie-heavy uniform segments
            if kind == 10: vals = rng.integers(300, 700, n); w = rng.integers(30, 301, n); cap = 4500   # sweep-like
            if kind == 11: vals = rng.integers(30000, 32768, n); cap = int(rng.integers(1, 600))        # top-cell check at the limit
            cases.append((vals, w, cap))
    got, st = _knapsack_batch(cases)
    assert not st[: len(cases)].any()
    for i, (vals, w, cap) in enumerate(cases):
        assert got[i] == c_oracle.knapsack(vals, w, cap), (i, i // 6)


def test_dp16_and_32bit_kernels_agree_on_a_dataset(monkeypatch):
    from summarizer_b200.batch import VideoBatch
    vids, scores = _dataset_videos("tvsum", 16)
    sc = torch.from_numpy(np.concatenate(scores))
    b = VideoBatch(vids)
    b.select(sc); torch.cuda.synchronize()
    a = b.picked.clone()
    monkeypatch.setenv("SMZ_NO_DP16", "1")
    b.picked.zero_()
    b.select(sc); torch.cuda.synchronize()
    assert torch.equal(a, b.picked)


@pytest.mark.parametrize("fused", [True, False])
def test_evaluate_fused_tail_equals_select_then_fscore(fused, monkeypatch):
    if not fused:
        monkeypatch.setenv("SMZ_NO_FUSED_TAIL", "1")
    from summarizer_b200.batch import VideoBatch
    vids, scores = _dataset_videos("tvsum", 13)
    sc = torch.from_numpy(np.concatenate(scores)).cuda()
    b = VideoBatch(vids)
    b.select(sc).fscore(); torch.cuda.synchronize()
    names = ("picked", "summary", "mask", "msum", "overlap", "gsum", "f", "avg_f", "max_f", "status")
    ref = {k: getattr(b, k).clone() for k in names}
    for k in names:
        getattr(b, k).fill_(0 if k != "status" else 7)
    b.evaluate(sc); b.check_status(); torch.cuda.synchronize()
    for k in names:
        assert torch.equal(ref[k], getattr(b, k)), k
    # annotator rows as bits
    h_bits = b.pack_user_summary_host(b.d_users.cpu(), n_threads=2)
    for k in ("overlap", "gsum", "f", "avg_f", "max_f"):
        getattr(b, k).zero_()
    b.evaluate(sc, d_bits=h_bits.cuda()); torch.cuda.synchronize()
    for k in names:
        assert torch.equal(ref[k], getattr(b, k)), k

"""Trainer.test / main.train on the device against the CPU oracle evaluated on the same model scores."""
import os
import numpy as np
import pytest
import torch
from scipy import stats

from oracle import eval_np as E
from summarizer_b200.main import train
from summarizer_b200.utils.config import HParameters

pytestmark = pytest.mark.gpu


def make_hps(tmp_path, **kw):
    hps = HParameters()
    hps.log_root = str(tmp_path)
    hps.tensorboard = False
    args = dict(model="vasnet", use_cuda="yes", splits_files="summe", log_level="error", extra_params={})
    args.update(kw)
    hps.load_from_args(args)
    return hps


def oracle_test_metrics(trainer, keys, scores, proportion, method):
    corrs, avg_f, max_f = [], [], []
    for key, s in zip(keys, scores):
        d = trainer.dataset[key]
        n_frames = int(d["n_frames"][()])
        s = s.cpu().numpy()
        frame = E.upsample(s, n_frames, d["picks"][...])
        us = d["user_scores"][...]
        corrs.append(np.mean([stats.spearmanr(stats.rankdata(-frame), stats.rankdata(-us[i]))[0] for i in range(us.shape[0])]))
        summary = E.generate_summary(s, d["change_points"][...], n_frames, d["n_frame_per_seg"][...].tolist(),
                                     d["picks"][...], proportion, method)
        a, m = E.evaluate_summary(summary, d["user_summary"][...])
        avg_f.append(a); max_f.append(m)
    return np.mean(corrs), np.mean(avg_f), np.mean(max_f)


@pytest.mark.parametrize("splits,method", [("summe", "knapsack"), ("tvsum", "rank")])
def test_trainer_test_matches_oracle(tmp_path, splits, method):
    hps = make_hps(tmp_path, splits_files=splits, selection_algorithm=method)
    t = hps.model_class(hps, hps.splits_files[0]).reset()
    _, test_keys = t._get_train_test_keys(1)
    avg_corr, (avg_f, max_f) = t.test(1)
    t.model.eval()
    scores = t._score_keys(test_keys)
    c, a, m = oracle_test_metrics(t, test_keys, scores, hps.summary_proportion, method)
    assert avg_corr == pytest.approx(c, abs=1e-10)
    assert float(avg_f) == pytest.approx(float(a), rel=1e-6) and float(max_f) == pytest.approx(float(m), rel=1e-6)


def test_cross_validation_run_end_to_end(tmp_path):
    hps = make_hps(tmp_path, splits_files="overfit", epochs=3, test_every_epochs=1, lr=1e-4)
    results = train(hps)
    assert [r[0].split("/")[-1] for r in results] == ["tvsum_splits_overfit.json", "summe_splits_overfit.json"]
    for _, corr, avg_f, max_f in results:
        assert np.isfinite([corr, avg_f, max_f]).all() and 0 <= avg_f <= max_f <= 1
    import os
    from oracle import c_oracle
    from summarizer_b200.models import h5_file, open_dataset
    for sf in hps.splits_files:
        assert os.path.exists(hps.weights_path[sf])
        # <split>_preds.h5 (models/__init__.py:142-177): group = dataset file name, per key the four fields; the stored
        # machine summary is what the oracle's generate_summary makes of the stored scores
        ds = open_dataset(hps.dataset_of_file[sf])
        with h5_file(hps.pred_path[sf], "r") as f:
            g = f[os.path.basename(hps.dataset_of_file[sf])]
            assert sorted(g.keys()) == sorted(ds.keys())
            for key in list(ds.keys())[:6]:
                k, d = g[key], ds[key]
                assert sorted(k.keys()) == ["machine_scores", "machine_summary", "scores", "user_summary"]
                sc = k["scores"][...]
                assert sc.dtype == np.float32 and sc.shape == (d["picks"][...].shape[0],)
                assert np.array_equal(k["user_summary"][...], d["user_summary"][...])
                ref_sum, _ = c_oracle.generate_summary(sc, d["change_points"][...], int(d["n_frames"][()]),
                                                       d["n_frame_per_seg"][...], d["picks"][...])
                assert np.array_equal(k["machine_summary"][...], ref_sum)
                assert k["machine_scores"][...].shape == (int(d["n_frames"][()]),)
    sd = torch.load(hps.weights_path[hps.splits_files[0]])
    assert "attention_head_projection.weight" in sd and sd["k2.weight"].shape == (1, 1024)


def test_graph_replayed_training_steps_match_eager(tmp_path, monkeypatch):
    """From the second visit of a video its optimizer step is replayed as a CUDA graph (models/__init__.py
    _train_supervised): same kernels and arithmetic.  Dropout is switched off (its random stream differs between the
    two modes).  Adam turns the float atomics' summation-order noise into +-lr moves of the parameters whose gradient
    is ~0, so weights are compared through what they compute: the test-video scores after 3 epochs, against an
    eager-vs-eager control run that measures that noise."""
    import random
    from summarizer_b200.models import vasnet_autograd
    monkeypatch.setattr(vasnet_autograd, "draw_keep_masks", lambda *a, **k: None)
    scores, moved = {}, {}
    for tag, mode in (("eager", "no"), ("eager2", "no"), ("graph", "yes")):
        hps = make_hps(tmp_path / tag, splits_files="splits/summe_splits_overfit.json", epochs=3, test_every_epochs=5, lr=1e-4,
                       extra_params={"cuda_graphs": mode})
        torch.manual_seed(3); random.seed(3)
        t = hps.model_class(hps, hps.splits_files[0]).reset()
        init = {k: v.detach().clone() for k, v in t.model.state_dict().items()}
        res = t.train(0)
        assert np.isfinite(res).all()
        moved[tag] = max(float((init[k] - v).abs().max()) for k, v in t.model.state_dict().items())
        t.model.eval()
        scores[tag] = torch.cat(t._score_keys(t._get_train_test_keys(0)[1]))
    assert moved["graph"] > 1e-4 and moved["eager"] > 1e-4                     # both really trained
    noise = float((scores["eager"] - scores["eager2"]).abs().max())
    diff = float((scores["eager"] - scores["graph"]).abs().max())
    print(f"graph vs eager: max score difference {diff:.3e}; eager vs eager control {noise:.3e}")
    assert diff <= 3 * noise + 1e-4, (diff, noise)


@pytest.mark.parametrize("family", ["vasnet", "dsn", "slstm"])
def test_weight_copies_follow_fused_optimizer_steps(family):
    """The kernels read bf16 / packed COPIES of the parameters.  torch's fused Adam updates parameters in place without
    bumping Tensor._version, so a version-keyed cache would keep serving the old copies: after an optimizer step both
    the next training forward and the next inference call must see the new weights (= what a fresh module loaded with
    the same state dict computes)."""
    import copy
    torch.manual_seed(5)
    if family == "vasnet":
        from summarizer_b200.models.vasnet import VASNet as cls
    elif family == "dsn":
        from summarizer_b200.models.dsn import DSN as cls
    else:
        from summarizer_b200.models.sumgan import sLSTM as cls
    m = cls().cuda().eval()                       # eval(): no dropout, so the training-path forward is deterministic
    x = torch.rand(24, 1, 1024, device="cuda")
    x = x / x.norm(dim=2, keepdim=True)
    opt = torch.optim.Adam(m.parameters(), lr=3e-3, fused=True)
    with torch.no_grad():
        y_before = m(x).clone()
    for _ in range(2):
        opt.zero_grad()
        m(x).square().mean().backward()
        opt.step()
    with torch.no_grad():
        y_inf = m(x).clone()
    y_train = m(x).detach().clone()
    fresh = cls().cuda().eval()
    fresh.load_state_dict(copy.deepcopy(m.state_dict()))
    with torch.no_grad():
        y_ref = fresh(x)
    assert float((y_before - y_ref).abs().max()) > 1e-3            # the steps really changed the function
    assert torch.allclose(y_inf, y_ref, atol=1e-6), float((y_inf - y_ref).abs().max())
    assert torch.allclose(y_train, fresh(x).detach(), atol=1e-6)


def test_positional_embedding_does_not_accumulate_in_cached_features(tmp_path):
    """--max_pos: VASNet adds the positional embedding in place to its input (reference quirk); the trainer hands it a
    private copy so the device-resident features stay what the dataset holds."""
    hps = make_hps(tmp_path, splits_files="splits/summe_splits_overfit.json", epochs=2, test_every_epochs=1,
                   extra_params={"max_pos": "2000", "pos_embed": "simple"})
    t = hps.model_class(hps, hps.splits_files[0]).reset()
    keys, _ = t._get_train_test_keys(0)
    before = t._video_tensors(keys[0])[0].clone()
    res = t.train(0)
    assert np.isfinite(res).all()
    assert torch.equal(before, t._video_tensors(keys[0])[0])


def test_reset_reuses_tensors_and_step_graphs_across_folds(tmp_path):
    """reset() between folds re-initialises the weights INTO the existing tensors (same RNG stream as a fresh model), so
    the CUDA graphs captured for the per-video training steps and the optimizer survive from fold to fold; the best
    weights of the previous fold are detached first (the reference's best_weights aliases the live parameters)."""
    hps = make_hps(tmp_path, splits_files="summe", epochs=4, test_every_epochs=2, lr=1e-4)
    t = hps.model_class(hps, hps.splits_files[0])
    torch.manual_seed(7)
    t.reset()
    ptrs = [p.data_ptr() for p in t.model.parameters()]
    r0 = t.train(0)
    graphs0 = t._step_graphs
    assert graphs0 is not None and len(graphs0.graphs) > 0
    best0 = {k: v.clone() for k, v in t.best_weights.items()}
    torch.manual_seed(11)
    t.reset()
    assert [p.data_ptr() for p in t.model.parameters()] == ptrs                  # same tensors ...
    torch.manual_seed(11)
    fresh = hps.model_class(hps, hps.splits_files[0]).reset()
    for a, b in zip(t.model.state_dict().values(), fresh.model.state_dict().values()):
        assert torch.equal(a, b)                                                 # ... holding a fresh initialisation
    for k, v in t.best_weights.items():
        assert torch.equal(v, best0[k])                                          # fold 0's best weights survived the reset
    n_before = len(graphs0.graphs)
    r1 = t.train(1)
    assert t._step_graphs is graphs0 and len(graphs0.graphs) >= n_before         # graphs reused (new videos add theirs)
    assert all(torch.count_nonzero(st["exp_avg"]) > 0 for st in t.optimizer.state.values())
    assert np.isfinite(list(r0) + list(r1)).all()
    # a fold trained on reused graphs behaves like one trained by a fresh trainer: same initial weights, same data ->
    # the first epoch's loss agrees up to the dropout draw
    avg_corr, (avg_f, max_f) = t.test(1)
    assert 0 <= avg_f <= max_f <= 1 and -1 <= avg_corr <= 1


@pytest.mark.gpu
def test_concurrent_folds_train_side_by_side(tmp_path):
    """--concurrent_folds 3: the five folds of a split file train on three worker threads / streams of one
    GPU (eager first visit, graph capture on the worker's stream, replay) — every fold reports, the best fold's weights
    are written and load back, and the metrics are in the range the sequential loop gives."""
    from summarizer_b200.main import train
    from summarizer_b200.utils.config import HParameters

    def run(extra):
        hps = HParameters()
        hps.log_root, hps.tensorboard = str(tmp_path / ("c" if extra else "s")), False
        hps.load_from_args(dict(model="vasnet", use_cuda="yes", splits_files="splits/summe_splits.json", log_level="error",
                                epochs=4, test_every_epochs=1, extra_params=extra))
        (res,) = train(hps)
        sf = hps.splits_files[0]
        assert os.path.exists(hps.weights_path[sf]) and os.path.exists(hps.pred_path[sf])
        sd = torch.load(hps.weights_path[sf])
        assert "k1.weight" in sd and all(torch.isfinite(v).all() for v in sd.values())
        return np.asarray(res[1:], dtype=np.float64)

    seq = run({})
    con = run({"concurrent_folds": 3})
    assert np.isfinite(con).all() and 0 <= con[1] <= con[2] <= 1
    # different initial weights / key orders per run: same ballpark, not equality
    assert abs(con[1] - seq[1]) < 0.1 and abs(con[2] - seq[2]) < 0.1, (seq, con)


@pytest.mark.gpu
def test_concurrent_folds_dsn(tmp_path):
    """concurrent_folds with the REINFORCE trainer: the episode draws come from per-trainer device-side Philox state, so
    one fold capturing its step graphs does not block another fold's eager draws."""
    hps = HParameters()
    hps.log_root, hps.tensorboard = str(tmp_path), False
    hps.load_from_args(dict(model="dsn", use_cuda="yes", splits_files="splits/summe_splits.json", log_level="error",
                            epochs=3, test_every_epochs=1, extra_params={"concurrent_folds": 2}))
    (res,) = train(hps)
    assert np.isfinite(res[1:]).all() and 0 <= res[2] <= res[3] <= 1
    assert os.path.exists(hps.weights_path[hps.splits_files[0]])

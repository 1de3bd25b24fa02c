"""GPU: the trainers against golden TRAINING trajectories of the reference's own trainers
(tests/golden/train_golden.npz, produced by oracle/gen_golden_train.py from the unmodified VASNetTrainer.train,
DSNTrainer.train, SumGANTrainer.pretrain / .train — models/vasnet.py:171-238, dsn.py:60-183, sumgan.py:320-533).

Same tiny dataset (written to a real HDF5 file with the built-in writer and read back through the Trainer), same
seeds for the initial weights and the key order, and the reference's random draws replayed one for one: VASNet dropout
keep-masks, the DSN episodes' actions, every SumGAN noise tensor.  Compared: per-epoch / per-step losses, the six
SumGAN log terms, the VAE pre-training loss, DSN rewards (1e-2 relative, the bf16 bar of BASELINE.json), the metrics
Trainer.test returns after every epoch, and the Adam updates of sampled parameter entries."""
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import gen_golden_train as G
from summarizer_b200 import synthetic
from summarizer_b200.utils.config import HParameters

pytestmark = pytest.mark.gpu
GOLDEN = np.load(G.GOLDEN)
REL = 1e-2
# Spearman correlation of ~45 frame scores against the annotator scores: a rank statistic — two scores that differ in the
# 4th digit swap ranks under bf16 and move it by a few 1e-2 although every score is within 1e-2 of the reference's
CORR_ABS = 5e-2


def make_trainer(tmp_path, model, name, **more_extra):
    run = G.RUNS[name]
    ds = synthetic.ArrayDataset(G.tiny_videos(run["frames"]), name="summe")
    h5 = synthetic.write_dataset_h5(ds, str(tmp_path / "summarizer_dataset_summe_tiny.h5"))
    sf = str(tmp_path / "summe_tiny_splits.json")
    with open(sf, "w") as fh:
        json.dump([G.tiny_split(run["frames"])], fh)
    hps = HParameters()
    hps.log_root, hps.tensorboard, hps.datasets = str(tmp_path), False, [h5]
    extra = dict(run["extra"])
    extra["cuda_graphs"] = "no"                  # the replayed draws come from the host, step by step
    extra.update(more_extra)
    hps.load_from_args(dict(model=model, use_cuda="yes", splits_files=sf, log_level="error", epochs=run["epochs"],
                            lr=run["lr"], weight_decay=G.WEIGHT_DECAY, test_every_epochs=1, extra_params=extra))
    hps.writer = G.Recorder()
    torch.manual_seed(run["seed"])
    t = hps.model_class(hps, hps.splits_files[0]).reset()
    return t, hps, run


def check_scalars(name, hps, tags, rel=REL, abs_tol=0.0):
    for tag in tags:
        want = GOLDEN[f"{name}/{tag}"]
        got = np.asarray(hps.writer.scalars[tag])
        assert got.shape == want.shape, (tag, got, want)
        np.testing.assert_allclose(got, want, rtol=rel, atol=abs_tol, err_msg=f"{name}/{tag}")


def check_updates(name, model, before, min_cos, max_rel):
    """Adam updates (final - initial) of the sampled entries, all parameters together."""
    after = G.sampled_params(model)
    got = np.concatenate([after[k] - before[k] for k in sorted(before)])
    want = np.concatenate([GOLDEN[f"{name}/w1/{k}"] - GOLDEN[f"{name}/w0/{k}"] for k in sorted(before)])
    w0 = np.concatenate([GOLDEN[f"{name}/w0/{k}"] for k in sorted(before)])
    np.testing.assert_allclose(np.concatenate([before[k] for k in sorted(before)]), w0, rtol=0, atol=0)   # same initial weights
    cos = float(got @ want / (np.linalg.norm(got) * np.linalg.norm(want)))
    rel = float(np.linalg.norm(got - want) / np.linalg.norm(want))
    assert cos >= min_cos and rel <= max_rel, f"{name}: parameter updates cos {cos:.4f} rel {rel:.3f}"
    return cos, rel


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_vasnet_trainer_follows_the_reference_trajectory(tmp_path, monkeypatch, precision):
    """precision="fp32" (split-bf16 operands, --precision fp32): the same trajectory at the float32 bar —
    losses to 1e-4, Adam updates to 2 % (Adam's first steps are sign-like: lr * g / (|g| + eps), so entries whose
    gradient is ~0 amplify any difference)."""
    from summarizer_b200.models import vasnet_autograd
    t, hps, run = make_trainer(tmp_path, "vasnet", "vasnet", precision=precision)
    assert t.model.precision == precision
    step = {"n": 0}

    def replay_masks(lengths, device, generator=None):
        (T,) = lengths
        att, y, h = G.vasnet_keep_masks(step["n"], T)
        step["n"] += 1
        return att.reshape(-1).to(device), y.to(device), h.to(device)

    monkeypatch.setattr(vasnet_autograd, "draw_keep_masks", replay_masks)
    before = G.sampled_params(t.model)
    random.seed(run["seed"])
    ret = t.train(0)
    assert step["n"] == run["epochs"] * 2
    check_scalars("vasnet", hps, ["Train/Loss"], rel=REL if precision == "bf16" else 1e-4)
    check_scalars("vasnet", hps, ["Test/Correlation"], rel=0, abs_tol=CORR_ABS if precision == "bf16" else 5e-3)
    check_scalars("vasnet", hps, ["Test/F-score_avg", "Test/F-score_max"], rel=1e-6)
    np.testing.assert_allclose(np.asarray(ret, dtype=np.float64)[1:], GOLDEN["vasnet/return"][1:], rtol=1e-6)
    cos, rel = check_updates("vasnet", t.model, before, *((0.99, 0.15) if precision == "bf16" else (0.9995, 0.03)))
    print(f"vasnet ({precision}): loss {hps.writer.scalars['Train/Loss']} vs {GOLDEN['vasnet/Train/Loss']}; updates cos {cos:.4f} rel {rel:.3f}")


def test_dsn_trainer_follows_the_reference_trajectory(tmp_path, monkeypatch):
    from summarizer_b200.models.dsn import DSNTrainer
    t, hps, run = make_trainer(tmp_path, "dsn", "dsn")
    acts, lens = GOLDEN["dsn/actions"], GOLDEN["dsn/action_lengths"]
    offs = np.concatenate([[0], np.cumsum(lens)])
    state = {"n": 0}

    def replay_actions(self, probs):
        E, T = self.num_episodes, probs.shape[0]
        rows = []
        for e in range(E):
            k = state["n"]
            assert lens[k] == T
            rows.append(torch.from_numpy(acts[offs[k]:offs[k + 1]].astype(np.uint8)))
            state["n"] += 1
        return torch.stack(rows).reshape(E, T).to(probs.device)

    monkeypatch.setattr(DSNTrainer, "_draw_actions", replay_actions)
    before = G.sampled_params(t.model)
    random.seed(run["seed"])
    ret = t.train(0)
    assert state["n"] == len(lens)
    check_scalars("dsn", hps, ["Train/Loss", "Train/Reward"])
    check_scalars("dsn", hps, ["Test/Correlation"], rel=0, abs_tol=CORR_ABS)
    check_scalars("dsn", hps, ["Test/F-score_avg", "Test/F-score_max"], rel=1e-6)
    cos, rel = check_updates("dsn", t.model, before, min_cos=0.999, max_rel=0.03)
    print(f"dsn: loss {hps.writer.scalars['Train/Loss']} vs {GOLDEN['dsn/Train/Loss']}; updates cos {cos:.4f} rel {rel:.3f}")
    assert np.isfinite(ret).all()


class ReplayNoise:
    """The reference's draws by position: the generator numbers them in call order (pretrain: eps; selector/encoder:
    eps; decoder: eps, uniform, eps'; discriminator: eps, uniform, eps', then the three input-noise tensors while
    epoch < epoch_noise, sumgan.py:419-468)."""
    OFFSET = {"eps": 0, "uniform": 1, "eps_p": 2, "noise_x": 3, "noise_x_hat": 4, "noise_x_hat_p": 5}

    def __init__(self):
        self.base, self.used = 0, 0

    def begin(self, phase):
        self.base += self.used
        self.used = 0

    def _draw(self, kind, t, role):
        off = self.OFFSET[role]
        self.used = max(self.used, off + 1)
        return G.noise_tensor(self.base + off, kind, tuple(t.shape)).to(device=t.device, dtype=t.dtype)

    def randn_like(self, t, role):
        return self._draw("randn", t, role)

    def rand_like(self, t, role):
        return self._draw("rand", t, role)


@pytest.mark.parametrize("name", ["sumgan", "sumgan_sup"])
def test_sumgan_trainer_follows_the_reference_trajectory(tmp_path, monkeypatch, name):
    from summarizer_b200.models import sumgan
    t, hps, run = make_trainer(tmp_path, "sumgan", name)
    replay = ReplayNoise()
    monkeypatch.setattr(sumgan, "noise", replay)
    lines = []
    monkeypatch.setattr(t.log, "info", lambda msg, *a, **k: lines.append(str(msg)))
    before = G.sampled_params(t.model)
    random.seed(run["seed"])
    t.train(0)
    replay.begin("end")
    assert replay.base == int(GOLDEN[f"{name}/noise_draws"][0])           # every draw of the reference was replayed
    tags = ["Train/Lse", "Train/Ld", "Train/Lc", "Train/D_x", "Train/D_x_hat", "Train/D_x_hat_p"]
    check_scalars(name, hps, tags)
    check_scalars(name, hps, ["Test/Correlation"], rel=0, abs_tol=CORR_ABS)
    lvae = [float(l.split("Lvae:")[1]) for l in lines if "Lvae:" in l]
    np.testing.assert_allclose(lvae, GOLDEN[f"{name}/Lvae"], rtol=REL)
    cos, rel = check_updates(name, t.model, before, min_cos=0.999, max_rel=0.05)
    print(f"{name}: Lse {hps.writer.scalars['Train/Lse']} vs {GOLDEN[name + '/Train/Lse']}; updates cos {cos:.4f} rel {rel:.3f}")

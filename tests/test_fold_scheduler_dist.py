"""CPU, world_size 2 over gloo: main.train deals folds to ranks, gathers the per-fold results and returns the
same `results` as a single process (no data-path collective)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from summarizer_b200.main import train
from summarizer_b200.models import Trainer
from summarizer_b200.utils.config import HParameters


class StubTrainer(Trainer):
    """Deterministic per-fold results without a device (the scheduler under test does not care)."""

    def _init_model(self):
        return torch.nn.Linear(2, 1)

    def train(self, fold):
        self._get_train_test_keys(fold)
        with torch.no_grad():
            self.model.weight.fill_(float(fold))                   # the weights remember which fold produced them
        self.best_weights = self.model.state_dict()
        # best fold: 3 for SumMe; 1 for TVSum, tied with fold 4, which must lose against the earlier fold
        corr = {"summe": [0.1, 0.2, 0.3, 0.9, 0.5], "tvsum": [0.1, 0.9, 0.3, 0.4, 0.9]}[self.dataset_name][fold]
        return corr, 0.2 + 0.01 * fold, 0.5 + 0.01 * fold

    def predict_dataset(self, pred_path):
        open(pred_path, "wb").close()


def make_hps(root):
    hps = HParameters()
    hps.log_root, hps.tensorboard, hps.allow_cpu = root, False, True
    hps.load_from_args(dict(model="random", use_cuda="no", splits_files="splits/summe_splits.json,splits/tvsum_splits.json",
                            log_level="error", extra_params={}))
    hps.model_class = StubTrainer
    return hps


def _best_fold_of_saved_weights(hps):
    return [int(torch.load(hps.weights_path[sf])["weight"][0, 0].item()) for sf in hps.splits_files]


def _worker(rank, world, root, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hps = make_hps(root)
    res = train(hps)
    out[rank] = ([(os.path.basename(f), float(c), float(a), float(m)) for f, c, a, m in res], hps.log_path,
                 _best_fold_of_saved_weights(hps) if rank == 0 else None)
    dist.destroy_process_group()


def test_two_ranks_return_single_process_results(tmp_path):
    hps1 = make_hps(str(tmp_path / "single"))
    single = train(hps1)
    want = [(os.path.basename(f), float(c), float(a), float(m)) for f, c, a, m in single]
    assert [w[0] for w in want] == ["summe_splits.json", "tvsum_splits.json"]
    assert _best_fold_of_saved_weights(hps1) == [3, 1]             # first fold reaching the maximum (main.py:33-35)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, str(tmp_path / "two"), 29731, out), nprocs=2, join=True)
    assert out[0][0] == want and out[1][0] == want
    assert out[0][1] == out[1][1]                                  # ONE log directory for the run (rank 0's, broadcast)
    assert out[0][2] == [3, 1]                                     # the owning rank shipped the best fold's weights to rank 0


# ---- data-parallel single-split training: gradient all-reduce (gloo here, NCCL on the GPU box) -------------------
from summarizer_b200.models.logistic import LogisticRegression  # noqa: E402


class CpuLogisticTrainer(Trainer):
    """The shared supervised loop with a plain torch model on the CPU and a stubbed device evaluation."""

    def _init_model(self):
        torch.manual_seed(1234 + (dist.get_rank() if dist.is_initialized() else 0))    # replicas start DIFFERENT
        return LogisticRegression()

    def test(self, fold):
        return 0.0, (0.0, 0.0)

    def train(self, fold):
        return self._train_supervised(fold)


def make_dp_hps(root, data_parallel):
    hps = HParameters()
    hps.log_root, hps.tensorboard, hps.allow_cpu = root, False, True
    hps.load_from_args(dict(model="logistic", use_cuda="no", splits_files="splits/summe_splits_overfit.json", log_level="error",
                            epochs=2, lr=1e-2, extra_params={"data_parallel": True} if data_parallel else {}))
    hps.model_class = CpuLogisticTrainer
    return hps


def _dp_worker(rank, world, root, port, out):
    import random
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    random.seed(7)
    hps = make_dp_hps(os.path.join(root, f"dp{rank}"), True)
    t = hps.model_class(hps, hps.splits_files[0]).reset()
    t.train(0)
    out[rank] = [p.detach().clone().numpy() for p in t.model.parameters()]
    dist.destroy_process_group()


def test_data_parallel_gradient_allreduce_matches_manual_averaging(tmp_path):
    import random
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(2, str(tmp_path), 29741, out), nprocs=2, join=True)
    for a, b in zip(out[0], out[1]):
        np.testing.assert_array_equal(a, b)                         # replicas stay in lock-step
    # single process: same initial weights (rank 0's), same key order, gradients of 2 videos averaged per step
    random.seed(7)
    hps = make_dp_hps(str(tmp_path / "ref"), False)
    t = hps.model_class(hps, hps.splits_files[0])
    torch.manual_seed(1234)
    t.model = LogisticRegression()
    keys, _ = t._get_train_test_keys(0)
    opt = torch.optim.Adam(t.model.parameters(), lr=hps.lr, weight_decay=hps.weight_decay)
    for _ in range(hps.epochs):
        random.shuffle(keys)
        for i in range(0, len(keys), 2):
            opt.zero_grad()
            for k in keys[i:i + 2]:
                seq, target = t._video_tensors(k)
                (torch.nn.functional.mse_loss(t.model(seq), target) / len(keys[i:i + 2])).backward()
            opt.step()
    for a, p in zip(out[0], t.model.parameters()):
        np.testing.assert_allclose(a, p.detach().numpy(), rtol=1e-5, atol=1e-7)


# ---- SumGANTrainer's per-phase update in data-parallel mode (BASELINE config 4): _groups + _update ----------------
def _sumgan_update_worker(rank, world, port, out):
    import random
    from summarizer_b200.models.sumgan import SumGANTrainer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class _H:
        lr, weight_decay, extra_params = 1e-2, 0.0, {"data_parallel": True}

    t = SumGANTrainer.__new__(SumGANTrainer)
    t.hps = _H
    torch.manual_seed(5)
    t.model = torch.nn.Linear(4, 3)
    data = {k: torch.randn(4, generator=torch.Generator().manual_seed(i)) for i, k in enumerate("abc")}
    opt = t._adam(t.model.parameters())
    dp, r, w = t._dp()
    assert dp is not None and (r, w) == (rank, world)
    random.seed(3)
    seen = []
    for key, n_active in t._groups(list("abc"), dp, r, w):         # 3 videos on 2 replicas: one idle slot
        seen.append((key, n_active))
        loss = None if key is None else 40.0 * t.model(data[key]).pow(2).sum()
        t._update(opt, loss, dp, n_active)
    out[rank] = ([p.detach().clone().numpy() for p in t.model.parameters()], seen)
    dist.destroy_process_group()


def test_sumgan_phase_update_data_parallel(tmp_path):
    import random
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sumgan_update_worker, args=(2, 29751, out), nprocs=2, join=True)
    (p0, seen0), (p1, seen1) = out[0], out[1]
    for a, b in zip(p0, p1):
        np.testing.assert_array_equal(a, b)
    assert [n for _, n in seen0] == [2, 1] and seen1[1][0] is None and seen0[1][0] is not None
    order = [seen0[0][0], seen1[0][0], seen0[1][0]]
    assert sorted(order) == list("abc")
    # single process: averaged gradients of each group, clip at 5.0, Adam
    torch.manual_seed(5)
    model = torch.nn.Linear(4, 3)
    data = {k: torch.randn(4, generator=torch.Generator().manual_seed(i)) for i, k in enumerate("abc")}
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, weight_decay=0.0)
    for group in (order[:2], order[2:]):
        opt.zero_grad()
        for k in group:
            (40.0 * model(data[k]).pow(2).sum() / len(group)).backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
        opt.step()
    for a, p in zip(p0, model.parameters()):
        np.testing.assert_allclose(a, p.detach().numpy(), rtol=1e-5, atol=1e-7)


def _sumgan_two_optimizer_worker(rank, world, port, out):
    """Two sub-networks with their own optimizers, gradients far above the clip threshold, a different video per rank:
    the stale gradients of the sub-network that is NOT stepping are rank-local and enter the clip norm."""
    from summarizer_b200.models.sumgan import SumGANTrainer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class _H:
        lr, weight_decay, extra_params = 1e-2, 0.0, {"data_parallel": True}

    t = SumGANTrainer.__new__(SumGANTrainer)
    t.hps = _H
    torch.manual_seed(11)
    t.model = torch.nn.ModuleDict({"a": torch.nn.Linear(4, 4), "b": torch.nn.Linear(4, 2)})
    opt_a, opt_b = t._adam(t.model["a"].parameters()), t._adam(t.model["b"].parameters())
    dp, r, w = t._dp()
    x = torch.randn(3, 4, generator=torch.Generator().manual_seed(100 + rank))      # every rank sees another video
    for step in range(4):
        for opt in (opt_a, opt_b):
            loss = 300.0 * t.model["b"](torch.relu(t.model["a"](x))).pow(2).sum()   # both sub-networks get gradients
            t._update(opt, loss, dp, world)
    out[rank] = [p.detach().clone().numpy() for p in t.model.parameters()]
    dist.destroy_process_group()


def test_sumgan_clip_is_shared_across_replicas_with_stale_local_gradients(tmp_path):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sumgan_two_optimizer_worker, args=(2, 29761, out), nprocs=2, join=True)
    for a, b in zip(out[0], out[1]):
        np.testing.assert_array_equal(a, b)

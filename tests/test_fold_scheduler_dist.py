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
        self.best_weights = self.model.state_dict()
        return 0.1 * (fold + 1), 0.2 + 0.01 * fold, 0.5 + 0.01 * fold

    def predict_dataset(self, pred_path):
        open(pred_path + ".npz", "wb").close()


def make_hps(root):
    hps = HParameters()
    hps.log_root, hps.tensorboard = root, False
    hps.load_from_args(dict(model="random", use_cuda="no", splits_files="summe", log_level="error", extra_params={}))
    hps.model_class = StubTrainer
    return hps


def _worker(rank, world, root, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = train(make_hps(os.path.join(root, f"r{rank}")))
    out[rank] = [(os.path.basename(f), float(c), float(a), float(m)) for f, c, a, m in res]
    dist.destroy_process_group()


def test_two_ranks_return_single_process_results(tmp_path):
    single = train(make_hps(str(tmp_path / "single")))
    want = [(os.path.basename(f), float(c), float(a), float(m)) for f, c, a, m in single]
    assert want[0][1] == np.mean([0.1 * (f + 1) for f in range(5)])
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, str(tmp_path), 29731, out), nprocs=2, join=True)
    assert out[0] == want and out[1] == want

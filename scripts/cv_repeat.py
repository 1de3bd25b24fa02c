"""Repeated in-process CV runs (bench.cv_stage's workload): wall time per pass, sequential then concurrent_folds=2."""
import gc, json, os, sys, tempfile, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from summarizer_b200 import main as M, synthetic
from summarizer_b200.utils.config import HParameters
cache = {n: synthetic.make_dataset(n) for n in ("tvsum", "summe")}
synthetic.make_dataset = lambda name, *a, **k: cache[name]
out = {}
for k in (1, 2):
    hps = HParameters()
    hps.load_from_args({"use_cuda": "yes", "cuda_device": 0, "model": "vasnet", "epochs": 20, "test_every_epochs": 10,
                        "splits_files": "splits/tvsum_splits.json,splits/summe_splits.json", "log_level": "error",
                        "log_root": tempfile.mkdtemp(prefix="smz_cv_"), "tensorboard": False,
                        "extra_params": {"concurrent_folds": k} if k > 1 else {}})
    walls = []
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        M.train(hps)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        gc.collect(); torch.cuda.synchronize(); t2 = time.perf_counter()
        walls.append((round(t1 - t0, 3), round(t2 - t1, 3), round(torch.cuda.memory_reserved() / 2**30, 2)))
    out[f"k{k}"] = walls
print(json.dumps(out))

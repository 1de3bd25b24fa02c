"""5-fold CV x 2 datasets in-process: sequential folds vs concurrent_folds=K on one GPU, at a given epoch count."""
import json, os, sys, tempfile, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from summarizer_b200 import main as M, synthetic
from summarizer_b200.utils.config import HParameters
epochs = int(os.environ.get("EPOCHS", 60))
model = os.environ.get("MODEL", "vasnet")
cache = {n: synthetic.make_dataset(n) for n in ("tvsum", "summe")}
synthetic.make_dataset = lambda name, *a, **k: cache[name]
out = {"epochs": epochs, "model": model}
for k in [1] + [int(x) for x in os.environ.get("KS", "2,4").split(",")]:
    hps = HParameters()
    hps.load_from_args({"use_cuda": "yes", "cuda_device": 0, "model": model, "epochs": epochs, "test_every_epochs": max(epochs // 3, 1),
                        "splits_files": "splits/tvsum_splits.json,splits/summe_splits.json", "log_level": "error",
                        "log_root": tempfile.mkdtemp(prefix="smz_cv_"), "tensorboard": False,
                        "extra_params": {"concurrent_folds": k} if k > 1 else {}})
    walls = []
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = M.train(hps)
        torch.cuda.synchronize(); walls.append(time.perf_counter() - t0)
    out[f"k{k}"] = {"wall_s": walls, "cv": [[float(x) for x in r[1:]] for r in res]}
print(json.dumps(out))

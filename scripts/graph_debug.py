import os, sys, random, tempfile, traceback
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from summarizer_b200.utils.config import HParameters
from summarizer_b200.main import train

for splits, graphs in (("splits/tvsum_splits_overfit.json", "yes"), ("overfit", "no"), ("overfit", "yes")):
    hps = HParameters(); hps.log_root = tempfile.mkdtemp(); hps.tensorboard = False
    hps.load_from_args(dict(model="vasnet", use_cuda="yes", splits_files=splits, log_level="error",
                            epochs=3, test_every_epochs=1, lr=1e-4, extra_params={"cuda_graphs": graphs}))
    try:
        res = train(hps)
        print(splits, graphs, "ok", [(r[0].split("/")[-1], round(float(r[1]), 4), round(float(r[2]), 4)) for r in res], flush=True)
    except Exception as e:
        traceback.print_exc()
        print(splits, graphs, "FAILED", type(e).__name__, e, flush=True)

"""Where the in-process fold-parallel CV spends its time (development aid): python scripts/dev/cv_time.py  [under torchrun]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import bench
from summarizer_b200 import main as M
import summarizer_b200.models as MM
orig_train = MM.Trainer.train
T = {}
def wrap(name, cls, attr):
    f = getattr(cls, attr)
    def g(self, *a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(self, *a, **k)
        torch.cuda.synchronize(); T[name] = T.get(name, 0) + time.perf_counter() - t0
        return r
    setattr(cls, attr, g)
wrap("test", MM.Trainer, "test")
wrap("predict_dataset", MM.Trainer, "predict_dataset")
wrap("train_supervised(incl test)", MM.Trainer, "_train_supervised")
wrap("reset", MM.Trainer, "reset")
wrap("draw_gtscores", MM.Trainer, "draw_gtscores")
wrap("draw_scores", MM.Trainer, "draw_scores")
for i in range(2):
    T.clear()
    out = bench.cv_stage(dev, int(os.environ.get("RANK", 0)), world)
    print("rank", os.environ.get("RANK", 0), "pass", i, "wall", round(out["wall_s"], 3), {k: round(v, 3) for k, v in T.items()}, flush=True)
if world > 1:
    dist.destroy_process_group()

"""torchrun target: --data_parallel VASNet training of one split with the step (incl. the NCCL all-reduce) replayed as CUDA
graphs vs eager: wall time, and that every replica ends with the same weights and reports the same fold results."""
import json, os, sys, tempfile, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from summarizer_b200 import main as M, synthetic
from summarizer_b200.utils.config import HParameters
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cache = {n: synthetic.make_dataset(n) for n in ("tvsum", "summe")}
synthetic.make_dataset = lambda name, *a, **k: cache[name]
out = {"world": world}
for mode in ("no", "yes"):
    hps = HParameters()
    hps.load_from_args({"use_cuda": "yes", "cuda_device": local, "model": "vasnet", "epochs": int(os.environ.get("EPOCHS", 12)),
                        "test_every_epochs": 6, "splits_files": "splits/tvsum_splits.json", "log_level": "error",
                        "log_root": tempfile.mkdtemp(prefix=f"smz_dp_{rank}_"), "tensorboard": False,
                        "extra_params": {"data_parallel": True, "dp_cuda_graphs": mode}})
    torch.manual_seed(1)
    import random; random.seed(1)
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    res = M.train(hps)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    sd = torch.load(hps.weights_path[hps.splits_files[0]]) if rank == 0 else None
    mine = [float(x) for x in res[0][1:]]
    allres = [None] * world
    dist.all_gather_object(allres, mine)
    out[f"graphs_{mode}"] = {"wall_s": dt, "results": mine, "replicas_report_the_same": all(r == allres[0] for r in allres)}
if rank == 0:
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()

"""torchrun target: only bench.train_dp_stage (data-parallel VASNet step, eager and graph-replayed)."""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
out = bench.train_dp_stage(dev, rank, world, dist)
if rank == 0:
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()

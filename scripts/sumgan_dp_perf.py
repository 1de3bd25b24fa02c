"""BASELINE config 4: SumGAN three-phase training, data-parallel over the ranks of one box (one video per rank and step,
NCCL all-reduce of each phase's gradients).  torchrun --nproc-per-node N scripts/sumgan_dp_perf.py"""
import json, os, sys, time
import torch
import torch.distributed as dist
import torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.gen_golden_models import make_input
from summarizer_b200.models.sumgan import SumGAN, SumGANTrainer

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    saved = os.dup(1); os.dup2(2, 1)                 # NCCL's banner goes to stderr
    dist.init_process_group("nccl"); dist.barrier()
    os.dup2(saved, 1); os.close(saved)


class _H:
    lr, weight_decay = 5e-5, 1e-5
    extra_params = {"data_parallel": world > 1}


torch.manual_seed(0)
m = SumGAN().to(dev).train()
t = SumGANTrainer.__new__(SumGANTrainer)
t.model, t.hps, t.sup, t.sigma, t.epoch_noise = m, _H, False, 0.3, 0
dp, r, w = t._dp()
if dp is not None:
    t._dp_sync_model(dp)
t.s_e_optimizer = t._adam(list(m.summarizer.s_lstm.parameters()) + list(m.summarizer.vae.e_lstm.parameters()))
t.d_optimizer = t._adam(m.summarizer.vae.d_lstm.parameters())
t.c_optimizer = t._adam(m.gan.c_lstm.parameters())
t.loss_BCE = nn.BCELoss()
T = 320
x = make_input(100 + rank, T, 1).to(dev)
y = torch.rand(T, 1, 1, device=dev)
for _ in range(2):
    t.train_step(x, y, 1, dp, world)
torch.cuda.synchronize()
if dp is not None:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    t.train_step(x, y, 1, dp, world)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
if dp is not None:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
chk = torch.stack([p.detach().double().sum() for p in m.parameters()]).sum()
same = True
if dp is not None:
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(lo == hi)
if rank == 0:
    print(json.dumps({"config": "SumGAN three-phase training, data-parallel, T=320 per rank", "n_gpus": world, "ms_per_step": ms.item(),
                      "frames_per_s": world * T / ms.item() * 1e3, "videos_per_step": world, "replicas_identical": same,
                      "allreduce_floats_per_step": sum(p.numel() for p in m.parameters())}))
if dp is not None:
    dist.destroy_process_group()

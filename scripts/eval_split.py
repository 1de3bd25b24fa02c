"""Selection + F-score stages on a sweep-shaped batch (per-kernel split comes from ncu gpu__time_duration)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from summarizer_b200 import synthetic
V = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda")
batch = synthetic.make_sweep_batch(V, dev, seed=5000)
g = torch.Generator(device=dev); g.manual_seed(1)
scores = torch.rand(batch.total_scores, generator=g, device=dev)
for _ in range(2):
    batch.select(scores); batch.fscore()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ts, tf = [], []
for _ in range(reps):
    ev[0].record(); batch.select(scores); ev[1].record(); batch.fscore(); ev[2].record()
    torch.cuda.synchronize()
    ts.append(ev[0].elapsed_time(ev[1])); tf.append(ev[1].elapsed_time(ev[2]))
batch.check_status()
print(json.dumps(dict(videos=V, select_ms=min(ts), fscore_ms=min(tf), avg_f_sum=float(batch.avg_f[:V].sum()),
                      picked_sum=int(batch.picked.sum()), msum=int(batch.msum[:V].sum()))))

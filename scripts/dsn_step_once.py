"""One DSN REINFORCE step + one evaluate_scores call (for ncu captures of the LSTM / reward / rank kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.distributions import Bernoulli
from summarizer_b200 import synthetic
from summarizer_b200.models.dsn import DSN, compute_rewards
from summarizer_b200.rankcorr import CorrBatch
torch.manual_seed(0)
v = synthetic.make_video("tvsum", 1)
x = torch.from_numpy(v["features"]).cuda().unsqueeze(1)
m = DSN().cuda().train()
for _ in range(2):
    probs = m(x); dist = Bernoulli(probs)
    actions = torch.stack([dist.sample() for _ in range(5)])
    rewards = compute_rewards(x, actions.reshape(5, -1))
    loss = -sum(dist.log_prob(actions[e]).mean() * rewards[e] for e in range(5)) / 5
    for p in m.parameters(): p.grad = None
    loss.backward()
cb = CorrBatch([(int(v["n_frames"]), v["user_scores"])])
frame_scores = torch.rand(int(v["n_frames"]), device="cuda")
print(float(cb.correlate(frame_scores)[0]), float(loss))

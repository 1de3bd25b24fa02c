"""Three eager VASNet training steps on one T=720 video (for an ncu launch list of the step's kernels)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from summarizer_b200.models import clip_grad_norm_, make_adam
from summarizer_b200.models.vasnet import VASNet
dev = torch.device("cuda")
T = int(os.environ.get("T", 720))
g = torch.Generator(device=dev); g.manual_seed(3)
x = torch.randn(T, 1, 1024, generator=g, device=dev).abs_()
x = x / x.norm(dim=2, keepdim=True)
t = torch.rand(T, 1, 1, generator=g, device=dev)
torch.manual_seed(0)
vas = VASNet().to(dev).train()
opt = make_adam(vas.parameters(), 5e-5, 1e-5)
for i in range(3):
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.mse_loss(vas(x), t)
    loss.backward(); opt.step()
    torch.cuda.synchronize()
    if i == 1:
        torch.cuda.nvtx.range_push("step") if False else None
print(float(loss))

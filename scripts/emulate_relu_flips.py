import sys; sys.path.insert(0,'/root/repo')
import torch, math
import torch.nn.functional as F
from oracle.gen_golden_models import build_vasnet, make_input
from summarizer_b200.models.vasnet import VASNet
torch.manual_seed(0)
m = build_vasnet(VASNet, 21, {}, 6.0)
x = make_input(40, 300, 1)[:,0]
g = torch.Generator().manual_seed(9); t = torch.rand(300, generator=g)
def run(round_k1, bias_shift=0.0):
    sd = {k: v.detach().clone().requires_grad_(True) for k,v in m.state_dict().items()}
    bf = lambda a: (a.bfloat16().float() - a).detach() + a   # straight-through rounding
    K = x @ sd["K.weight"].t(); Q = x @ sd["Q.weight"].t(); V = x @ sd["V.weight"].t()
    e = (Q @ K.t()) * m.scale
    al = torch.softmax(e, 1)
    y = (al @ V) @ sd["attention_head_projection.weight"].t() + x
    yn = F.layer_norm(y, (1024,), sd["layer_norm.weight"], sd["layer_norm.bias"], m.epsilon)
    W1 = sd["k1.weight"]
    if round_k1: yn, W1 = bf(yn), bf(W1)
    h = torch.relu(yn @ W1.t() + sd["k1.bias"] + bias_shift)
    hn = F.layer_norm(h, (1024,), sd["layer_norm.weight"], sd["layer_norm.bias"], m.epsilon)
    s = torch.sigmoid(hn @ sd["k2.weight"].t() + sd["k2.bias"]).reshape(-1)
    loss = ((s - t)**2).mean(); loss.backward()
    return {k: v.grad for k,v in sd.items()}
for shift in (0.0, 6.0):
    a = run(False, shift); b = run(True, shift)
    print("bias shift", shift, {k: f"{((a[k]-b[k]).norm()/a[k].norm()).item():.2e}" for k in a})

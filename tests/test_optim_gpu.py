"""GPU: the library's Adam / clip_grad_norm_ kernels against torch.optim.Adam / torch.nn.utils.clip_grad_norm_
(the update rules of models/vasnet.py:160-161,211-212, models/dsn.py:100,147-149, models/sumgan.py:268-275,433-436)."""
import pytest
import torch

from summarizer_b200 import optim as O

pytestmark = pytest.mark.gpu
SHAPES = [(1024, 1024), (1024,), (1,), (3, 5, 7), (4099,), (2048, 1024), (17,)] + [(33,)] * 70     # > 64 tensors: two launches


def make(seed):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(*s, generator=g, device="cuda")) for s in SHAPES]


def set_grads(ps, seed, scale=1.0, skip=()):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    for i, p in enumerate(ps):
        p.grad = None if i in skip else scale * torch.randn(p.shape, generator=g, device="cuda")


@pytest.mark.parametrize("wd", [0.0, 1e-5, 0.1])
def test_adam_matches_torch(wd):
    a, b = make(0), make(0)
    oa = O.Adam(a, lr=5e-3, weight_decay=wd)
    ob = torch.optim.Adam(b, lr=5e-3, weight_decay=wd)
    for step in range(6):
        skip = (3,) if step == 2 else ()
        set_grads(a, 100 + step, skip=skip); set_grads(b, 100 + step, skip=skip)
        oa.step(); ob.step()
        for x, y in zip(a, b):
            assert torch.allclose(x, y, rtol=1e-5, atol=1e-6), (step, x.shape, (x - y).abs().max().item())
    sa, sb = oa.state_dict()["state"], ob.state_dict()["state"]
    assert sa.keys() == sb.keys()
    for k in sa:
        assert set(sa[k].keys()) == {"step", "exp_avg", "exp_avg_sq"}
        assert float(sa[k]["step"]) == float(sb[k]["step"])          # per-parameter counters (parameter 3 skipped one step)
        assert torch.allclose(sa[k]["exp_avg"], sb[k]["exp_avg"], rtol=1e-5, atol=1e-7)
        assert torch.allclose(sa[k]["exp_avg_sq"], sb[k]["exp_avg_sq"], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("scale,max_norm", [(1.0, 5.0), (1e-4, 5.0), (30.0, 0.5)])
def test_clip_grad_norm_matches_torch(scale, max_norm):
    a, b = make(1), make(1)
    set_grads(a, 7, scale, skip=(1,)); set_grads(b, 7, scale, skip=(1,))
    na = O.clip_grad_norm_(a, max_norm)
    nb = torch.nn.utils.clip_grad_norm_(b, max_norm)
    assert torch.allclose(na, nb, rtol=1e-5)
    for x, y in zip(a, b):
        assert (x.grad is None) == (y.grad is None)
        if x.grad is not None:
            assert torch.allclose(x.grad, y.grad, rtol=2e-5, atol=0)
    # bit-stable: the same gradients give the same norm bits
    set_grads(a, 7, scale, skip=(1,))
    assert torch.equal(O.clip_grad_norm_(a, max_norm), na)


def test_adam_step_replays_in_a_cuda_graph():
    a, b = make(2)[:4], make(2)[:4]
    oa, ob = O.Adam(a, lr=1e-2, weight_decay=1e-5), torch.optim.Adam(b, lr=1e-2, weight_decay=1e-5)
    set_grads(a, 5); set_grads(b, 5)
    oa.step(); ob.step()                      # state created outside the capture
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        O.clip_grad_norm_(a, 5.0)
        oa.step()
    for _ in range(3):
        g.replay()
    for _ in range(3):                        # the same three (clip, step) pairs eagerly with torch on the twin
        torch.nn.utils.clip_grad_norm_(b, 5.0)
        ob.step()
    torch.cuda.synchronize()
    for x, y in zip(a, b):
        assert torch.allclose(x, y, rtol=1e-5, atol=1e-6), (x - y).abs().max().item()
    assert all(float(st["step"]) == 4.0 for st in oa.state.values())        # the device-side counters advanced with every replay


def test_rejects_non_float32_or_cpu():
    p = torch.nn.Parameter(torch.zeros(4, device="cuda", dtype=torch.float64))
    p.grad = torch.zeros_like(p)
    with pytest.raises(Exception):
        O.Adam([p]).step()
    with pytest.raises(Exception):
        O.clip_grad_norm_([p], 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 45, 1000, 4097])
def test_mse_loss_matches_torch(n):
    """optim.mse_loss (smz_mse_loss: loss and gradient in one launch) vs torch.nn.MSELoss() and its autograd."""
    from summarizer_b200.optim import mse_loss
    g = torch.Generator(device="cuda"); g.manual_seed(n)
    s = torch.rand(n, 1, 1, generator=g, device="cuda", requires_grad=True)
    t = torch.rand(n, 1, 1, generator=g, device="cuda")
    s_ref = s.detach().clone().requires_grad_(True)
    loss = mse_loss(s, t)
    ref = torch.nn.MSELoss()(s_ref, t)
    (3.0 * loss).backward(); (3.0 * ref).backward()
    assert loss.shape == ref.shape == ()
    assert torch.allclose(loss, ref, rtol=1e-6, atol=0)
    assert torch.allclose(s.grad, s_ref.grad, rtol=1e-6, atol=1e-12)
    # shapes torch would broadcast, or a target that needs a gradient: torch's implementation
    t2 = t.clone().requires_grad_(True)
    assert torch.allclose(mse_loss(s.detach(), t2), torch.nn.functional.mse_loss(s.detach(), t2))

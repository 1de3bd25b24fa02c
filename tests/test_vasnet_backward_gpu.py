"""VASNet training step (smz_vasnet_forward(training) + smz_vasnet_backward through torch.autograd) against
torch autograd over the float32 restatement of the reference forward (oracle/models_torch.py), with the SAME
dropout keep-masks on both sides.  bf16 tensor-core path, loss within 1e-2 relative.

Gradient tolerance.  A ReLU sits behind the k1 GEMM (vasnet.py:140-141): rounding that GEMM's operands to bf16
flips the sign of ~0.2 % of the pre-activations that lie within rounding distance of 0, and every flipped unit
switches its gradient entry on or off.  That alone moves each gradient by 4-5 % in relative L2 norm (reproduced
on the CPU by rounding ONLY k1's operands in the float32 oracle, scripts/emulate_relu_flips.py) although the loss
is unchanged to 1e-3.  So every case runs twice: with k1.bias shifted by +6 (all units active, gradients smooth)
the bar is 3e-2 per parameter (typically 2e-3 .. 1e-2; the scalar k2.bias gradient is a heavily cancelling sum) — this
is the check of every backward formula — and with the reference's
initialisation the bar is 0.12 relative L2 and cosine similarity > 0.99."""
import numpy as np
import pytest
import torch

from oracle import models_torch as MT
from oracle.gen_golden_models import build_vasnet, make_input
from summarizer_b200.models.vasnet import VASNet
from summarizer_b200.models.vasnet_autograd import draw_keep_masks, vasnet_apply

pytestmark = pytest.mark.gpu

GRAD_TOL_SMOOTH, GRAD_TOL, COS_MIN = 3e-2, 0.12, 0.99
NAMES = ["Q.weight", "K.weight", "V.weight", "attention_head_projection.weight", "k1.weight", "k1.bias", "k2.weight",
         "k2.bias", "layer_norm.weight", "layer_norm.bias"]


def oracle_loss_and_grads(m, xs, targets, masks):
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items() if k in NAMES}
    outs, o_att, o_row = [], 0, 0
    for x in xs:
        T = x.shape[0]
        kw = {}
        if masks is not None:
            kw = dict(keep_att=masks[0][o_att:o_att + T * T].view(T, T).float(), keep_y=masks[1][o_row:o_row + T].float(),
                      keep_h=masks[2][o_row:o_row + T].float())
        outs.append(MT.vasnet_forward(sd, x, scale=m.scale, eps=m.epsilon, aperture=m.aperture, ignore_self=m.ignore_self, **kw))
        o_att += T * T; o_row += T
    s = torch.cat(outs)
    loss = ((s - targets) ** 2).mean()
    loss.backward()
    return loss.item(), s.detach(), {k: v.grad for k, v in sd.items()}


@pytest.mark.parametrize("bias_shift", [6.0, 0.0], ids=["all_active", "reference_init"])
@pytest.mark.parametrize("lengths,kw,dropout", [([300], {}, False), ([130], {}, True), ([64], {"attention_aperture": 9}, True),
                                                ([50, 77], {"ignore_self": True}, True), ([707], {}, False)])
def test_gradients_match_oracle(lengths, kw, dropout, bias_shift):
    m = build_vasnet(VASNet, 21, kw, 6.0).cuda()
    with torch.no_grad():
        m.k1.bias.add_(bias_shift)
    m.train(dropout)
    xs = [make_input(40 + i, T, 1)[:, 0].cuda() for i, T in enumerate(lengths)]
    g = torch.Generator(device="cuda"); g.manual_seed(9)
    masks = draw_keep_masks(lengths, xs[0].device, g) if dropout else None
    targets = torch.rand(sum(lengths), generator=g, device="cuda")
    scores = vasnet_apply(m, torch.cat(xs), lengths, masks=masks)
    loss = ((scores - targets) ** 2).mean()
    loss.backward()
    ref_loss, ref_scores, ref_g = oracle_loss_and_grads(m, xs, targets, masks)
    assert abs(loss.item() - ref_loss) / ref_loss < 1e-2
    assert (scores.detach() - ref_scores).abs().max().item() < 1e-2
    errs = {name: (p.grad - ref_g[name]).norm().item() / max(ref_g[name].norm().item(), 1e-12)
            for name, p in m.named_parameters() if name in ref_g}
    cos = {name: torch.nn.functional.cosine_similarity(p.grad.flatten(), ref_g[name].flatten(), dim=0).item()
           for name, p in m.named_parameters() if name in ref_g}
    report = ", ".join(f"{k} {v:.2e}" for k, v in errs.items())
    print("relative gradient errors:", report)
    assert max(errs.values()) < (GRAD_TOL_SMOOTH if bias_shift else GRAD_TOL), report
    assert min(cos.values()) > COS_MIN, cos


def test_module_forward_trains_like_the_reference_loop():
    """nn.Module path: model(seq) -> MSELoss -> backward -> Adam (vasnet.py:207-212), loss goes down."""
    torch.manual_seed(0)
    m = VASNet().cuda().train()
    opt = torch.optim.Adam(m.parameters(), lr=5e-5, weight_decay=1e-5)
    x = make_input(3, 200, 1).cuda()
    target = torch.linspace(0, 1, 200, device="cuda").view(200, 1, 1)
    losses = []
    for _ in range(30):
        y = m(x)
        loss = torch.nn.functional.mse_loss(y, target)
        opt.zero_grad(); loss.backward(); opt.step()
        losses.append(float(loss))
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < np.mean(losses[:5])


def test_input_gradient_with_learned_positional_embedding():
    m = build_vasnet(VASNet, 22, {"max_length": 64, "pos_embed": "simple"}, 0.5).cuda().eval()
    x = make_input(5, 48, 1).cuda()
    y = m(x.clone())
    y.sum().backward()
    assert m.pos_embed.weight.grad is not None and torch.isfinite(m.pos_embed.weight.grad).all()
    # oracle: same forward in float32 torch with the embedding added outside
    emb = m.pos_embed.weight.detach().clone().requires_grad_(True)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    s = MT.vasnet_forward(sd, x[:, 0] + emb[:48], scale=m.scale, eps=m.epsilon)
    s.sum().backward()
    err = (m.pos_embed.weight.grad - emb.grad).norm().item() / emb.grad.norm().item()
    assert err < 3e-2, f"positional-embedding gradient error {err:.3e}"


# float32-accurate mode: every contraction of the forward AND the backward on split-bf16 operands.  With all ReLU units
# active the gradients agree with the float32 oracle to 1e-5 .. 6e-5 (measured; asserted at 2e-4).  With the reference initialisation the k1
# pre-activations are within ~1e-5 of the oracle's, so only the handful of units (of ~3 * 10^5) that sit closer to 0 than
# that take the other side of the ReLU; a gradient is discontinuous there — f flipped units of n move it by ~sqrt(f / n) in
# relative L2 norm (3 of 3 * 10^5: 3e-3; the bf16 mode's 0.2 %: the 4-9 % of the header) — which any two float32
# implementations with different summation orders show as well.  The bar drops from 0.12 to 1.5e-2.
GRAD_TOL_FP32_SMOOTH, GRAD_TOL_FP32 = 2e-4, 1.5e-2


@pytest.mark.parametrize("bias_shift", [6.0, 0.0], ids=["all_active", "reference_init"])
@pytest.mark.parametrize("lengths,kw,dropout", [([300], {}, False), ([130], {}, True), ([64], {"attention_aperture": 9}, True),
                                                ([50, 77], {"ignore_self": True}, True), ([707], {}, False)])
def test_fp32_mode_gradients_match_oracle(lengths, kw, dropout, bias_shift):
    m = build_vasnet(VASNet, 21, kw, 6.0).cuda()
    with torch.no_grad():
        m.k1.bias.add_(bias_shift)
    m.precision = "fp32"
    m.train(dropout)
    xs = [make_input(40 + i, T, 1)[:, 0].cuda() for i, T in enumerate(lengths)]
    g = torch.Generator(device="cuda"); g.manual_seed(9)
    masks = draw_keep_masks(lengths, xs[0].device, g) if dropout else None
    targets = torch.rand(sum(lengths), generator=g, device="cuda")
    scores = vasnet_apply(m, torch.cat(xs), lengths, masks=masks)
    loss = ((scores - targets) ** 2).mean()
    loss.backward()
    ref_loss, ref_scores, ref_g = oracle_loss_and_grads(m, xs, targets, masks)
    assert abs(loss.item() - ref_loss) / ref_loss < 1e-4
    rel = ((scores.detach() - ref_scores).abs() / ref_scores.abs().clamp_min(1e-6)).max().item()
    assert rel < 1e-4, rel
    errs = {name: (p.grad - ref_g[name]).norm().item() / max(ref_g[name].norm().item(), 1e-12)
            for name, p in m.named_parameters() if name in ref_g}
    report = ", ".join(f"{k} {v:.2e}" for k, v in errs.items())
    print("fp32 mode relative gradient errors:", report)
    assert max(errs.values()) < (GRAD_TOL_FP32_SMOOTH if bias_shift else GRAD_TOL_FP32), report


def test_fp32_mode_input_gradient():
    m = build_vasnet(VASNet, 22, {"max_length": 64, "pos_embed": "simple"}, 0.5).cuda().eval()
    m.precision = "fp32"
    x = make_input(5, 48, 1).cuda()
    m(x.clone()).sum().backward()
    emb = m.pos_embed.weight.detach().clone().requires_grad_(True)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    MT.vasnet_forward(sd, x[:, 0] + emb[:48], scale=m.scale, eps=m.epsilon).sum().backward()
    err = (m.pos_embed.weight.grad - emb.grad).norm().item() / emb.grad.norm().item()
    assert err < GRAD_TOL_FP32, f"positional-embedding gradient error {err:.3e}"


def test_library_keep_masks_are_fair_bits_and_advance_on_the_device():
    """smz_dropout_keep_masks (the module's dropout draws): 0/1 bytes, P(keep) = 1/2, independent between positions and
    calls; the call number is bumped on the device (graph replays draw fresh masks); a seed reproduces the stream."""
    from summarizer_b200.models.vasnet_autograd import mask_state
    lengths = [130, 77]
    st = mask_state(torch.device("cuda"), seed=99)
    a = draw_keep_masks(lengths, torch.device("cuda"), st)
    assert st.tolist() == [99, 1, 0]
    b = draw_keep_masks(lengths, torch.device("cuda"), st)
    assert st.tolist() == [99, 2, 0]
    assert a[0].shape == (130 * 130 + 77 * 77,) and a[1].shape == (207, 1024) and a[2].shape == (207, 1024)
    for m in (*a, *b):
        assert m.dtype == torch.uint8 and int(m.max()) == 1 and int(m.min()) == 0
    flat_a, flat_b = torch.cat([m.flatten() for m in a]).float(), torch.cat([m.flatten() for m in b]).float()
    n = flat_a.numel()
    for f in (flat_a, flat_b):
        assert abs(float(f.mean()) - 0.5) < 4 * 0.5 / n ** 0.5
    agree = float((flat_a == flat_b).float().mean())                   # independent calls agree on half the positions
    assert abs(agree - 0.5) < 4 * 0.5 / n ** 0.5
    lag = float((flat_a[1:] == flat_a[:-1]).float().mean())            # neighbouring bits are independent
    assert abs(lag - 0.5) < 4 * 0.5 / n ** 0.5
    again = draw_keep_masks(lengths, torch.device("cuda"), mask_state(torch.device("cuda"), seed=99))
    assert all(torch.equal(x, y) for x, y in zip(a, again))
    # tails that are not a multiple of 128 bytes / unaligned outputs
    from summarizer_b200 import _native as N
    buf = torch.full((1000,), 7, dtype=torch.uint8, device="cuda")
    N.check(N.lib().smz_dropout_keep_masks(N.ptr(st), buf[3:].data_ptr(), 900, N.current_stream()))
    assert set(buf[3:903].tolist()) == {0, 1} and set(buf[903:].tolist()) == {7} and set(buf[:3].tolist()) == {7}

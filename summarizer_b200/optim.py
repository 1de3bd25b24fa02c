"""Optimizer step of the training loops on the library's kernels (csrc/smz_optim.cu):

* ``Adam`` — ``torch.optim.Adam`` semantics (L2 ``weight_decay`` added to the gradient, bias corrections, ``eps`` outside
  the square root; ``amsgrad`` / ``maximize`` are not offered) with the same ``state_dict`` layout (``step``, ``exp_avg``,
  ``exp_avg_sq`` per parameter), so optimizer checkpoints interchange with the reference's
  (models/vasnet.py:160-161, models/dsn.py:100, models/sumgan.py:268-275).  One kernel launch per 64 parameter tensors
  plus a one-thread counter bump; the step counter lives on the device, so a captured training step replays correctly.
* ``clip_grad_norm_`` — ``torch.nn.utils.clip_grad_norm_(parameters, max_norm)`` for the L2 norm (dsn.py:147,
  sumgan.py:433-436): a bit-stable two-stage sum of squares and one in-place scaling launch; returns the total norm as a
  0-d device tensor (no host synchronisation).

float32 CUDA parameters only — anything else raises (there is no CPU fallback).
"""
import ctypes as C

import torch

from . import _native as N


class OptimTensor(C.Structure):
    """struct smz_optim_tensor (include/summarizer_b200.h)."""
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("step", C.c_void_p), ("n", C.c_int64)]


def _check(p):
    if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
        raise N.NativeError("summarizer_b200.optim works on contiguous float32 CUDA tensors only")


def _table(rows):
    arr = (OptimTensor * max(len(rows), 1))()
    for i, (p, g, m, v, t, n) in enumerate(rows):
        arr[i] = OptimTensor(p, g, m, v, t, n)
    return arr


def clip_grad_norm_(parameters, max_norm):
    """In-place gradient clipping by the total L2 norm; returns the norm (0-d float32 device tensor)."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    ps = [p for p in parameters if p.grad is not None]
    if not ps:
        return torch.zeros(())
    rows = []
    for p in ps:
        g = p.grad
        _check(g)
        rows.append((None, g.data_ptr(), None, None, None, g.numel()))
    dev = ps[0].device
    tab = _table(rows)
    need = C.c_int64(0)
    N.check(N.lib().smz_grad_sqnorm_workspace_floats(tab, len(rows), C.byref(need)))
    # scratch per call (the caching allocator makes it free; inside a CUDA-graph capture it must come from the graph's pool)
    ws = torch.empty(max(int(need.value), 1), dtype=torch.float32, device=dev)
    sq = torch.empty(1, dtype=torch.float32, device=dev)
    st = N.current_stream()
    N.check(N.lib().smz_grad_sqnorm(tab, len(rows), N.ptr(sq), N.ptr(ws), ws.numel(), st))
    N.check(N.lib().smz_clip_grads(tab, len(rows), N.ptr(sq), C.c_float(float(max_norm)), st))
    return sq.sqrt().reshape(())


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        st = N.current_stream()
        for group in self.param_groups:
            rows = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                _check(p); _check(p.grad)
                s = self.state[p]
                if not s:
                    s["step"] = torch.zeros((), dtype=torch.float32, device=p.device)      # on the device, as torch's capturable Adam
                    s["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    s["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                rows.append((p.data_ptr(), p.grad.data_ptr(), s["exp_avg"].data_ptr(), s["exp_avg_sq"].data_ptr(),
                             s["step"].data_ptr(), p.numel()))
            if not rows:
                continue
            b1, b2 = group["betas"]
            N.check(N.lib().smz_adam_step(_table(rows), len(rows), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                          float(group["weight_decay"]), st))
        return loss

    def zero_state_(self):
        """A fresh optimizer in place (state tensors and step counters zeroed, addresses kept): what a new fold needs when
        the captured step graphs of the previous folds are to be replayed."""
        for s in self.state.values():
            for k in ("step", "exp_avg", "exp_avg_sq"):
                if k in s:
                    s[k].zero_()


class _MseLoss(torch.autograd.Function):
    """mean((scores - target)^2) and its gradient w.r.t. the scores in one launch (smz_mse_loss)."""

    @staticmethod
    def forward(ctx, scores, target):
        s = scores.detach().reshape(-1).contiguous()
        t = target.detach().reshape(-1).contiguous()
        loss = torch.empty((), dtype=torch.float32, device=s.device)
        ds = torch.empty_like(s)
        N.check(N.lib().smz_mse_loss(N.ptr(s), N.ptr(t), s.numel(), N.ptr(loss), N.ptr(ds), N.current_stream()))
        ctx.save_for_backward(ds)
        ctx.shape = scores.shape
        return loss

    @staticmethod
    def backward(ctx, grad):
        (ds,) = ctx.saved_tensors
        return (ds * grad).reshape(ctx.shape), None


def mse_loss(scores, target):
    """``torch.nn.MSELoss()(scores, target)`` (the supervised trainers' criterion, vasnet.py:199,208) on the library's kernel
    for float32 CUDA tensors of equal shape whose target needs no gradient; torch's implementation otherwise."""
    if (scores.is_cuda and scores.dtype == torch.float32 and target.dtype == torch.float32 and target.is_cuda
            and scores.shape == target.shape and not target.requires_grad and scores.numel() > 0):
        return _MseLoss.apply(scores, target)
    return torch.nn.functional.mse_loss(scores, target)

"""Dense contractions on the tcgen05 GEMM (smz_gemm_bf16) for the host-side composition of the LSTM family:
input projections, dX and weight gradients of whole sequences, and ``nn.Linear`` with autograd.  Replaces the
cuBLAS calls behind ``nn.Linear`` / ``nn.LSTM``'s input GEMMs in models/sumgan.py:58-59,84,112 — operands are
bfloat16, accumulation and outputs float32."""
import ctypes as C

import torch

from . import _native as N

OUT_F32, RELU, RES_F32, BIAS_M = 1, 2, 4, 8


def _check_operand(t, name):
    if t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: bfloat16 matrix with contiguous rows expected")
    if t.stride(0) % 8 or t.data_ptr() % 16:
        raise ValueError(f"{name}: leading dimension must be a multiple of 8 elements and the base 16-byte aligned")


def gemm(a, b, a_mn=False, b_mn=False, bias=None, out=None, accumulate=False, alpha=1.0):
    """float32 C[M,N] = alpha * A·Bᵀ (+ bias per column) (+ C when ``accumulate``).
    a is [M,K] (or [K,M] with a_mn), b is [N,K] (or [K,N] with b_mn): the storage forms of smz_gemm_bf16."""
    N.require_device()
    _check_operand(a, "a"); _check_operand(b, "b")
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    Nn, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if K != Kb:
        raise ValueError(f"gemm: inner dimensions differ ({K} vs {Kb})")
    if out is None:
        out = torch.empty(M, Nn, dtype=torch.float32, device=a.device)
        if accumulate:
            out.zero_()
    if K == 0:
        if not accumulate:
            out.zero_()
        return out
    flags = OUT_F32 | (RES_F32 if accumulate else 0)
    N.check(N.lib().smz_gemm_bf16(int(a_mn), int(b_mn), N.ptr(a), a.stride(0), N.ptr(b), b.stride(0), N.ptr(out),
                                  out.stride(0), M, Nn, K, float(alpha), N.ptr(bias), N.ptr(out) if accumulate else None,
                                  out.stride(0), flags, N.current_stream()))
    return out


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        xb = x.detach().to(torch.bfloat16).contiguous()
        wb = weight.detach().to(torch.bfloat16).contiguous()
        ctx.save_for_backward(xb, wb)
        ctx.has_bias = bias is not None
        return gemm(xb, wb, bias=None if bias is None else bias.detach().float().contiguous())

    @staticmethod
    def backward(ctx, dy):
        xb, wb = ctx.saved_tensors
        dyb = dy.to(torch.bfloat16).contiguous()
        dx = gemm(dyb, wb, b_mn=True) if ctx.needs_input_grad[0] else None
        dw = gemm(dyb, xb, a_mn=True, b_mn=True) if ctx.needs_input_grad[1] else None
        db = dy.sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db


def linear(x, weight, bias=None):
    """``F.linear`` for a 2-D float32 ``x`` [rows, in] with in/out multiples of 8, on the tcgen05 GEMM."""
    return _LinearFn.apply(x, weight, bias)

"""tcgen05 GEMM building block (smz_gemm_bf16_tn) against a plain PyTorch fp32 reference of the same op
(bf16-rounded operands, fp32 accumulation).  Runs on the B200 box only."""
import ctypes as C

import numpy as np
import pytest
import torch

from summarizer_b200 import _native as N

pytestmark = pytest.mark.gpu

OUT_F32, RELU, RES_F32, BIAS_M = 1, 2, 4, 8
OUT_F16, A_F16, B_F16 = 2048, 4096, 8192


def gemm(a, b, alpha=1.0, bias=None, residual=None, flags=0, ldc=None):
    M, K = a.shape
    Nn = b.shape[0]
    ldc = Nn if ldc is None else ldc
    out = torch.full((M, ldc), float("nan"), device=a.device,
                     dtype=torch.float32 if flags & OUT_F32 else (torch.float16 if flags & OUT_F16 else torch.bfloat16))
    N.check(N.lib().smz_gemm_bf16_tn(N.ptr(a), a.stride(0), N.ptr(b), b.stride(0), N.ptr(out), ldc, M, Nn, K,
                                     C.c_float(alpha), N.ptr(bias), N.ptr(residual),
                                     0 if residual is None else residual.stride(0), flags, N.current_stream()))
    return out[:, :Nn]


def ref(a, b, alpha=1.0, bias=None, residual=None, flags=0):
    r = alpha * (a.float() @ b.float().t())
    if bias is not None:
        r = r + (bias[:, None] if flags & BIAS_M else bias[None, :])
    if residual is not None:
        r = r + residual.float()
    if flags & RELU:
        r = torch.relu(r)
    return r


def describe_mismatch(got, want, tol):
    """Blind-debugging aid: where do the errors sit (rows / column blocks)?"""
    err = (got.float() - want).abs()
    bad = err > tol
    rows = torch.nonzero(bad.any(1)).flatten()[:8].tolist()
    cols = torch.nonzero(bad.any(0)).flatten()[:16].tolist()
    return (f"max err {err.max().item():.4g} (tol {tol:.3g}), bad {int(bad.sum())}/{bad.numel()}, "
            f"first bad rows {rows}, cols {cols}, nan {int(torch.isnan(got.float()).sum())}")


@pytest.mark.parametrize("M,Nn,K", [(128, 256, 64), (128, 256, 128), (128, 256, 1024), (256, 512, 1024),
                                    (300, 2048, 1024), (1024, 300, 1024), (77, 40, 64), (2000, 2000, 1024),
                                    (707, 1024, 768), (5000, 1024, 1024)])
def test_gemm_matches_torch(M, Nn, K):
    N.require_device()
    g = torch.Generator(device="cuda"); g.manual_seed(M * 7 + Nn * 3 + K)
    a = torch.randn(M, K, generator=g, device="cuda").bfloat16()
    b = torch.randn(Nn, K, generator=g, device="cuda").bfloat16()
    ldc = (Nn + 7) // 8 * 8
    got = gemm(a, b, flags=OUT_F32, ldc=ldc)
    want = ref(a, b)
    tol = 1e-3 * (K ** 0.5) + 1e-3
    assert torch.allclose(got, want, atol=tol, rtol=1e-3), describe_mismatch(got, want, tol)
    got16 = gemm(a, b, alpha=0.125, ldc=ldc)
    want16 = ref(a, b, alpha=0.125)
    assert torch.allclose(got16.float(), want16, atol=tol, rtol=1e-2), describe_mismatch(got16, want16, tol)


def test_gemm_epilogue_variants():
    N.require_device()
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    M, Nn, K = 333, 1024, 1024
    a = torch.randn(M, K, generator=g, device="cuda").bfloat16()
    b = (torch.randn(Nn, K, generator=g, device="cuda") / 32).bfloat16()
    bias_n = torch.randn(Nn, generator=g, device="cuda")
    bias_m = torch.randn(M, generator=g, device="cuda")
    res32 = torch.randn(M, Nn, generator=g, device="cuda")
    res16 = res32.bfloat16()
    for kw in (dict(bias=bias_n, flags=OUT_F32), dict(bias=bias_m, flags=OUT_F32 | BIAS_M),
               dict(bias=bias_n, residual=res32, flags=OUT_F32 | RES_F32), dict(residual=res16, flags=OUT_F32),
               dict(bias=bias_n, flags=OUT_F32 | RELU), dict(bias=bias_n, residual=res16, flags=RELU)):
        got = gemm(a, b, **kw)
        want = ref(a, b, **kw)
        tol = 2e-2 if not kw["flags"] & OUT_F32 else 2e-3
        assert torch.allclose(got.float(), want, atol=tol, rtol=1e-2), (kw["flags"], describe_mismatch(got, want, tol))


def test_gemm_submatrix_operands():
    """Operands that are column slices of wider arrays (the Q and K halves of the packed projection)."""
    N.require_device()
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    T = 450
    qk = torch.randn(T, 2048, generator=g, device="cuda").bfloat16()
    got = gemm(qk[:, :1024], qk[:, 1024:], alpha=1 / 32, flags=OUT_F32, ldc=456)
    want = ref(qk[:, :1024], qk[:, 1024:], alpha=1 / 32)
    assert torch.allclose(got, want, atol=2e-2, rtol=1e-3), describe_mismatch(got, want, 2e-2)


def gemm_general(a_mn, b_mn, a, b, M, Nn, K, flags=OUT_F32, residual=None):
    out = torch.full((M, Nn), float("nan"), device=a.device, dtype=torch.float32 if flags & OUT_F32 else torch.bfloat16)
    N.check(N.lib().smz_gemm_bf16(int(a_mn), int(b_mn), N.ptr(a), a.stride(0), N.ptr(b), b.stride(0), N.ptr(out), Nn,
                                  M, Nn, K, C.c_float(1.0), None, N.ptr(residual), Nn if residual is not None else 0,
                                  flags, N.current_stream()))
    return out


@pytest.mark.parametrize("a_mn,b_mn", [(0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,Nn,K", [(128, 256, 64), (128, 256, 256), (1024, 1024, 300), (300, 1024, 1024),
                                    (2048, 1024, 707), (707, 707, 1024), (1024, 600, 777)])
def test_gemm_mn_major_operands(a_mn, b_mn, M, Nn, K):
    """op(A) stored [K, M] and/or op(B) stored [K, N] (the backward-pass forms), K tails zero-filled by TMA."""
    N.require_device()
    g = torch.Generator(device="cuda"); g.manual_seed(M + 3 * Nn + 5 * K + a_mn + 2 * b_mn)
    pad8 = lambda n: (n + 7) // 8 * 8
    am = torch.randn(M, K, generator=g, device="cuda").bfloat16()
    bm = torch.randn(Nn, K, generator=g, device="cuda").bfloat16()
    if a_mn:
        a = torch.zeros(K, pad8(M), device="cuda", dtype=torch.bfloat16); a[:, :M] = am.t()
    else:
        a = torch.zeros(M, pad8(K), device="cuda", dtype=torch.bfloat16); a[:, :K] = am
    if b_mn:
        b = torch.zeros(K, pad8(Nn), device="cuda", dtype=torch.bfloat16); b[:, :Nn] = bm.t()
    else:
        b = torch.zeros(Nn, pad8(K), device="cuda", dtype=torch.bfloat16); b[:, :K] = bm
    want = am.float() @ bm.float().t()
    got = gemm_general(a_mn, b_mn, a, b, M, Nn, K)
    tol = 1e-3 * (K ** 0.5) + 1e-3
    assert torch.allclose(got, want, atol=tol, rtol=1e-3), describe_mismatch(got, want, tol)
    # accumulate into an existing float32 buffer (weight-gradient accumulation): C = A.B^T + C
    acc = torch.randn(M, Nn, generator=g, device="cuda")
    out = acc.clone()
    N.check(N.lib().smz_gemm_bf16(int(a_mn), int(b_mn), N.ptr(a), a.stride(0), N.ptr(b), b.stride(0), N.ptr(out), Nn,
                                  M, Nn, K, C.c_float(1.0), None, N.ptr(out), Nn, OUT_F32 | RES_F32, N.current_stream()))
    assert torch.allclose(out, want + acc, atol=tol, rtol=1e-3), describe_mismatch(out, want + acc, tol)


@pytest.mark.parametrize("a16,b16,out16", [(True, True, True), (False, False, True), (True, True, False)])
def test_gemm_float16_operands(a16, b16, out16):
    """tcgen05 kind::f16 with float16 operands (both operands in the same format: the hardware rejects a bf16 x f16 mix
    as an illegal instruction) and / or float16 output, which keeps 11 significant bits."""
    N.require_device()
    M, Nn, K = 515, 768, 320
    g = torch.Generator(device="cuda"); g.manual_seed(99)
    a = torch.randn(M, K, generator=g, device="cuda").to(torch.float16 if a16 else torch.bfloat16)
    b = torch.randn(Nn, K, generator=g, device="cuda").to(torch.float16 if b16 else torch.bfloat16)
    flags = (A_F16 if a16 else 0) | (B_F16 if b16 else 0)
    want = ref(a, b)
    got = gemm(a, b, flags=flags | OUT_F32)
    assert torch.allclose(got, want, rtol=0, atol=1e-3 * K ** 0.5), describe_mismatch(got, want, 1e-3 * K ** 0.5)
    got = gemm(a, b, flags=flags | (OUT_F16 if out16 else 0))
    assert got.dtype == (torch.float16 if out16 else torch.bfloat16)
    rel = 2.0 ** (-11 if out16 else -8)
    assert ((got.float() - want).abs() <= rel * want.abs() + 1e-3 * K ** 0.5).all()
    assert ((got.float() - want).abs() / want.abs().clamp_min(1.0)).max().item() < 1.01 * rel + 2e-3


def split_planes(x32):
    """hi + lo bf16 planes of a float32 matrix through smz_split_bf16_multi, checked against the definition."""
    hi = torch.empty_like(x32, dtype=torch.bfloat16)
    lo = torch.empty_like(x32, dtype=torch.bfloat16)
    src = (C.c_void_p * 1)(x32.data_ptr()); dh = (C.c_void_p * 1)(hi.data_ptr()); dl = (C.c_void_p * 1)(lo.data_ptr())
    cnt = (C.c_int64 * 1)(x32.numel())
    N.check(N.lib().smz_split_bf16_multi(src, dh, dl, cnt, 1, N.current_stream()))
    assert torch.equal(hi, x32.bfloat16()) and torch.equal(lo, (x32 - x32.bfloat16().float()).bfloat16())
    return hi, lo


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,Nn,K", [(128, 256, 64), (304, 1024, 1024), (712, 704, 712), (2000, 1024, 2000)])
def test_gemm_split_bf16_is_float32_accurate(a_mn, b_mn, M, Nn, K):
    """smz_gemm_bf16_split (hi.hi + lo.hi + hi.lo in one float32 accumulator) vs a float64 product of the float32
    operands: ~1e-5 of the result's scale where plain bf16 operands give ~3e-3; float32 output, then hi + lo output
    planes whose sum carries the same accuracy."""
    N.require_device()
    g = torch.Generator(device="cuda"); g.manual_seed(M + 3 * Nn + 7 * K + a_mn + 2 * b_mn)
    a32 = torch.randn((K, M) if a_mn else (M, K), generator=g, device="cuda")
    b32 = torch.randn((K, Nn) if b_mn else (Nn, K), generator=g, device="cuda")
    a32 = a32 * torch.rand(a32.shape[1], generator=g, device="cuda") * 3        # uneven column scales
    (ah, al), (bh, bl) = split_planes(a32), split_planes(b32)
    want = ((a32.double().t() if a_mn else a32.double()) @ (b32.double() if b_mn else b32.double().t()))
    # error measure: |C - want| against sum_k |a_ik| |b_kj| (what a per-product relative error is a fraction of)
    scale = ((a32.double().abs().t() if a_mn else a32.double().abs()) @ (b32.double().abs() if b_mn else b32.double().abs().t()))

    def run(flags, c, c_lo, a_lo=al, b_lo=bl):
        N.check(N.lib().smz_gemm_bf16_split(a_mn, b_mn, N.ptr(ah), N.ptr(a_lo), ah.stride(0), N.ptr(bh), N.ptr(b_lo), bh.stride(0),
                                            N.ptr(c), N.ptr(c_lo), c.stride(0), M, Nn, K, C.c_float(1.0), None, None, 0, flags,
                                            N.current_stream()))
    out = torch.full((M, Nn), float("nan"), device="cuda")
    run(OUT_F32, out, None)
    err = ((out.double() - want).abs() / scale).max().item()
    plain = torch.empty_like(out)
    run(OUT_F32, plain, None, a_lo=None, b_lo=None)
    err_plain = ((plain.double() - want).abs() / scale).max().item()
    print(f"split-bf16 max err / (|A|.|B|) = {err:.2e} (bf16 operands: {err_plain:.2e})")
    assert err < 1e-5 and err_plain > 20 * err, (err, err_plain)      # worst case per product: 3 * 2^-18 = 1.1e-5
    # one-sided: B exact in bf16 -> two passes
    out2 = torch.empty_like(out)
    run(OUT_F32, out2, None, b_lo=None)
    want2 = ((a32.double().t() if a_mn else a32.double()) @ (bh.double() if b_mn else bh.double().t()))
    assert ((out2.double() - want2).abs() / scale).max().item() < 1e-5
    # hi + lo output planes
    if Nn % 8 == 0:
        ch = torch.full((M, Nn), float("nan"), device="cuda", dtype=torch.bfloat16)
        cl = torch.full((M, Nn), float("nan"), device="cuda", dtype=torch.bfloat16)
        run(0, ch, cl)
        assert torch.equal(ch, out.bfloat16()), describe_mismatch(ch, out.bfloat16().float(), 0)
        rec = ch.double() + cl.double()
        assert (rec - out.double()).abs().max().item() <= 2.0 ** -16 * out.abs().max().item()

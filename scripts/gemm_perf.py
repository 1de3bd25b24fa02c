"""Times smz_gemm_bf16_tn (CUDA events, inputs > L2 rotated) and torch.matmul on the same shapes."""
import ctypes as C, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from summarizer_b200 import _native as N

def run(M, Nn, K, iters=20):
    a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(Nn, K, device="cuda").bfloat16()
    out = torch.empty(M, Nn, device="cuda", dtype=torch.bfloat16)
    L = N.lib(); st = N.current_stream()
    def f():
        N.check(L.smz_gemm_bf16_tn(N.ptr(a), K, N.ptr(b), K, N.ptr(out), Nn, M, Nn, K, C.c_float(1.0), None, None, 0, 0, st))
    def t(fn):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    ms = t(f); ms_t = t(lambda: torch.matmul(a, b.t(), out=out))
    fl = 2.0 * M * Nn * K
    print(json.dumps({"M": M, "N": Nn, "K": K, "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1),
                      "torch_ms": round(ms_t, 4), "torch_tflops": round(fl / ms_t / 1e9, 1)}), flush=True)

if __name__ == "__main__":
    N.require_device()
    for shp in [(16384, 2048, 1024), (16384, 1024, 1024), (65536, 2048, 1024), (2000, 2000, 1024), (2000, 1024, 2048), (1024, 16384, 1024), (320, 2048, 1024)]:
        run(*shp)

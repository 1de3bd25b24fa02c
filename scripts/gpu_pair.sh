set -x
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x 2>&1 | tail -12
SMZ_GEMM_PAIR=0 timeout 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 200 python scripts/gemm_perf.py 2>&1 | head -8

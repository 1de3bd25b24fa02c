set -x
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu 2>&1 | tail -40 > gpurun_out/gemm_pytest.log
cat gpurun_out/gemm_pytest.log
timeout 200 python scripts/gemm_perf.py > gpurun_out/gemm_perf.log 2>&1
cat gpurun_out/gemm_perf.log

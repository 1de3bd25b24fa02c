set -x
timeout 600 python -m pytest tests/test_vasnet_gpu.py tests/test_gemm_gpu.py -q -m gpu 2>&1 | tail -80 > gpurun_out/vasnet_pytest.log
cat gpurun_out/vasnet_pytest.log
timeout 300 python scripts/vasnet_perf.py > gpurun_out/vasnet_perf.log 2>&1
cat gpurun_out/vasnet_perf.log

set -x
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/all_pytest.log
cat gpurun_out/all_pytest.log
timeout 300 python scripts/vasnet_perf.py > gpurun_out/vasnet_perf.log 2>&1
cat gpurun_out/vasnet_perf.log
timeout 200 python scripts/gemm_perf.py > gpurun_out/gemm_perf.log 2>&1
cat gpurun_out/gemm_perf.log

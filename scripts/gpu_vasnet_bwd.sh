set -x
SMZ_DEBUG_SYNC=1 timeout 600 python -m pytest tests/test_vasnet_backward_gpu.py -q -m gpu -rA 2>&1 | grep -E "relative gradient errors|passed|failed|Error" | tail -60 > gpurun_out/vasnet_bwd_pytest.log
cat gpurun_out/vasnet_bwd_pytest.log

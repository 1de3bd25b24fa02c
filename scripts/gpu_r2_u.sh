#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_vasnet_gpu.py tests/test_vasnet_backward_gpu.py tests/test_dsn_gpu.py tests/test_train_golden_gpu.py -m gpu -q -s -x 2>&1 | tail -60 > gpurun_out/r2u_pytest.txt
tail -40 gpurun_out/r2u_pytest.txt

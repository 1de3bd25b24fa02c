#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gemm_gpu.py tests/test_vasnet_gpu.py tests/test_sweep_golden_gpu.py tests/test_vasnet_backward_gpu.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -25
python scripts/vasnet_perf.py 2>&1 | tail -5 | tee gpurun_out/r2j_vasnet_perf.jsonl
python scripts/vasnet_steps.py 2>&1 | tail -25 | tee gpurun_out/r2j_vasnet_steps.txt

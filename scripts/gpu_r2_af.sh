#!/bin/bash
mkdir -p gpurun_out
for pdl in 1 0; do
  for i in 1 2 3 4 5; do
    SMZ_PDL=$pdl timeout 300 python -m pytest tests/test_train_golden_gpu.py tests/test_trainer_gpu.py -m gpu -q -x 2>&1 > gpurun_out/r2af_run.txt
    echo "PDL=$pdl run $i: $(tail -1 gpurun_out/r2af_run.txt)"
    if grep -q failed gpurun_out/r2af_run.txt; then grep -E "Error|error|smz_|\.py:[0-9]+: in" gpurun_out/r2af_run.txt | head -30; cp gpurun_out/r2af_run.txt gpurun_out/r2af_fail_pdl$pdl.txt; fi
  done
done

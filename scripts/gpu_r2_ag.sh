#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
  timeout 300 python -m pytest tests -m gpu -q -x 2>&1 > gpurun_out/r2ag_run.txt
  echo "run $i: $(tail -1 gpurun_out/r2ag_run.txt)"
  if grep -q failed gpurun_out/r2ag_run.txt; then cp gpurun_out/r2ag_run.txt gpurun_out/r2ag_fail.txt; grep -n "Error\|error" gpurun_out/r2ag_fail.txt | head -20; break; fi
done

#!/bin/bash
mkdir -p gpurun_out
EPOCHS=60 timeout 900 python scripts/cv_concurrent_perf.py > gpurun_out/r2aa_cv60.json 2> gpurun_out/r2aa_cv60.err
cat gpurun_out/r2aa_cv60.json; tail -3 gpurun_out/r2aa_cv60.err

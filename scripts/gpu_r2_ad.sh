#!/bin/bash
# round 2 final: ncu --set full of the scoring-stage GEMMs (after the split-plane / PDL changes) and of the fused
# evaluation kernel (pooling + 16-bit DP + summary + F-score tail) at 2 000 sweep videos
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trainer_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 5 -c 5 -o gpurun_out/r02h_gemm_full -f python scripts/dev/prof_score.py > gpurun_out/r02h_ncu_gemm.log 2>&1
tail -2 gpurun_out/r02h_ncu_gemm.log
bash scripts/ncu_summary.sh gpurun_out/r02h_gemm_full.ncu-rep gpurun_out/r02h_ncu_gemm_kernels_summary.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dp16_kernel" -s 4 -c 2 -o gpurun_out/r02h_eval_full -f python scripts/eval_perf.py 2000 > gpurun_out/r02h_ncu_eval.log 2>&1
tail -2 gpurun_out/r02h_ncu_eval.log
bash scripts/ncu_summary.sh gpurun_out/r02h_eval_full.ncu-rep gpurun_out/r02h_ncu_eval_kernels_2000videos_summary.csv
ls -la gpurun_out/r02h_*

#!/bin/bash
# round 2, call D: ncu --set full of dp16_kernel (plain and with the fused evaluation tail) and pool_smem_kernel at 10 000 videos
set -x
mkdir -p gpurun_out
SMZ_NO_FUSED_TAIL=1 ncu --set full --clock-control none --import-source on -k regex:dp16_kernel -s 3 -c 1 -f -o gpurun_out/r2d_dp16_plain python scripts/eval_perf.py 10000 > /dev/null 2>gpurun_out/r2d_ncu1.err
ncu --set full --clock-control none --import-source on -k regex:dp16_kernel -s 12 -c 1 -f -o gpurun_out/r2d_dp16_fused python scripts/eval_perf.py 10000 > /dev/null 2>gpurun_out/r2d_ncu2.err
ncu --set full --clock-control none --import-source on -k regex:pool_smem_kernel -s 3 -c 1 -f -o gpurun_out/r2d_pool python scripts/eval_perf.py 10000 > /dev/null 2>gpurun_out/r2d_ncu3.err
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/r2d_ncu*.err

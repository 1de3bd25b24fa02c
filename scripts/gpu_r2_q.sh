#!/bin/bash
# round-2 profiles: launch list + DRAM bytes of a 64-video sweep step, ncu --set full of the four GEMM launches of a chunk,
# and of the evaluation kernels at 2000 videos
P=r02e
mkdir -p gpurun_out
SMZ_BENCH_VIDEOS=64 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${P}_launches_sweep64.csv python bench.py --steps 1 --warmup 1 --cpu-seconds 1 --e2e-videos 16 > gpurun_out/${P}_ncu_launch.log 2>&1
tail -2 gpurun_out/${P}_ncu_launch.log
NV=32 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel --launch-skip 8 -c 8 -o gpurun_out/${P}_gemm_full -f python scripts/dev/prof_score.py > gpurun_out/${P}_ncu_gemm.log 2>&1
tail -2 gpurun_out/${P}_ncu_gemm.log
ncu --set full --clock-control none --import-source on -k regex:"dp16_kernel|dp_kernel|pool_regular_kernel|fscore_kernel|pack_user_bits_kernel" --launch-skip 8 -c 12 -o gpurun_out/${P}_eval_full -f python scripts/eval_perf.py 2000 > gpurun_out/${P}_ncu_eval.log 2>&1
tail -2 gpurun_out/${P}_ncu_eval.log
ls -la gpurun_out/${P}_*

#!/bin/bash
# round-2 final: ncu launch list (time + DRAM bytes per launch) of a 64-video sweep step with the final code
P=r02i
mkdir -p gpurun_out
SMZ_BENCH_VIDEOS=64 timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${P}_launches_sweep64.csv python bench.py --steps 1 --warmup 1 --cpu-seconds 1 --e2e-videos 16 > gpurun_out/${P}_ncu_launch.log 2>&1
tail -2 gpurun_out/${P}_ncu_launch.log | cut -c1-300
ls -la gpurun_out/${P}_*

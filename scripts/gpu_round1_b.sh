set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1b_smoke.log 2>&1; cat gpurun_out/r1b_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.err; cat gpurun_out/r1b_bench.json; tail -5 gpurun_out/r1b_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1b_bench_ref.json 2> gpurun_out/r1b_bench_ref.err; cat gpurun_out/r1b_bench_ref.json
SMZ_BENCH_VIDEOS=64 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 1 --warmup 1 --cpu-seconds 1 --e2e-videos 16 > gpurun_out/r1b_ncu_launch.log 2>&1
SMZ_BENCH_VIDEOS=64 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel" -s 30 -c 6 -o gpurun_out/r1b_gemm_full -f python bench.py --steps 1 --warmup 1 --cpu-seconds 1 --e2e-videos 16 > gpurun_out/r1b_ncu_full.log 2>&1
SMZ_BENCH_VIDEOS=64 ncu --set full --clock-control none --import-source on -k regex:"softmax_kernel|layernorm_kernel|head_kernel|fscore_kernel|dp_kernel" -s 10 -c 6 -o gpurun_out/r1b_rows_full -f python bench.py --steps 1 --warmup 1 --cpu-seconds 1 --e2e-videos 16 > gpurun_out/r1b_ncu_full2.log 2>&1
tail -3 gpurun_out/r1b_ncu_full.log

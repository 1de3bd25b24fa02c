set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_ref.json 2> gpurun_out/r1_bench_ref.err
SMZ_BENCH_VIDEOS=2000 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 1 --e2e-videos 64 > gpurun_out/r1_ncu_launch.log 2>&1
SMZ_BENCH_VIDEOS=2000 ncu --set full --clock-control none --import-source on -k regex:"fscore_kernel|dp_kernel|pool_kernel|summary_kernel" -c 4 -o gpurun_out/r1_eval_full -f python bench.py --steps 1 --warmup 0 --cpu-seconds 1 --e2e-videos 64 > gpurun_out/r1_ncu_full.log 2>&1
tail -5 gpurun_out/r1_pytest.log; cat gpurun_out/r1_bench.json; cat gpurun_out/r1_bench_ref.json

set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r1i_bench.json 2> gpurun_out/r1i_bench.err; tail -3 gpurun_out/r1i_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1i_bench_ref.json 2> gpurun_out/r1i_bench_ref.err
SMZ_BENCH_VIDEOS=64 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r1i_launches.csv python bench.py --steps 1 --warmup 1 --cpu-seconds 1 --e2e-videos 16 > gpurun_out/r1i_ncu_launch.log 2>&1
SMZ_BENCH_VIDEOS=64 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel" -s 30 -c 7 -o gpurun_out/r1i_gemm_full -f python bench.py --steps 1 --warmup 1 --cpu-seconds 1 --e2e-videos 16 > gpurun_out/r1i_ncu_full.log 2>&1
SMZ_BENCH_VIDEOS=64 ncu --set full --clock-control none --import-source on -k regex:"lstm_fwd_kernel|lstm_bwd_kernel|dp_kernel|fscore_kernel|reward_rows|pool_kernel|rank_kernel" -c 10 -o gpurun_out/r1i_other_full -f python bench.py --steps 1 --warmup 1 --cpu-seconds 1 --e2e-videos 16 > gpurun_out/r1i_ncu_full2.log 2>&1
python -c "
import json; d=json.load(open('gpurun_out/r1i_bench.json')); print(d['value'], d['frames_per_s'], d['stages_ms'], d['roofline']['achieved'], d['roofline']['frac'], d['e2e'], d['train'], d['cpu_baseline']['value'], d['clocks'])"
cat gpurun_out/r1i_bench_ref.json | cut -c1-300

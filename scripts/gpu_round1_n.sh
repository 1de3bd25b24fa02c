set -x
P=${1:-r1n}
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > gpurun_out/${P}_pytest.log; cat gpurun_out/${P}_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; tail -3 gpurun_out/${P}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${P}_bench_ref.json 2> gpurun_out/${P}_bench_ref.err
SMZ_BENCH_VIDEOS=64 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 1 --warmup 1 --cpu-seconds 1 --e2e-videos 16 > gpurun_out/${P}_ncu_launch.log 2>&1
SMZ_BENCH_VIDEOS=64 ncu --set full --clock-control none --import-source on -k regex:"dp_kernel|fscore_kernel|fscore_bits_kernel|pool_smem_kernel|summary_kernel" -c 10 -o gpurun_out/${P}_eval_full -f python bench.py --steps 1 --warmup 1 --cpu-seconds 1 --e2e-videos 16 > gpurun_out/${P}_ncu_full2.log 2>&1
(cd summarizer_b200 && timeout 600 python benchmark.py -e 10 -c yes -s splits/summe_splits.json,splits/tvsum_splits.json > ../gpurun_out/${P}_benchmark_py.log 2>&1; tail -14 ../gpurun_out/${P}_benchmark_py.log)
python -c "
import json; d=json.load(open('gpurun_out/${P}_bench.json')); print(d['value'], d['frames_per_s'], d['stages_ms'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline_eval'], d['e2e'], d['train'], d['cpu_baseline']['value'], d['clocks'], d['gpu_launches'])"
cat gpurun_out/${P}_bench_ref.json | cut -c1-300

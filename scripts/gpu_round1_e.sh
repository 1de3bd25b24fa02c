set -x
timeout 600 python -m pytest tests/test_vasnet_gpu.py -q -m gpu 2>&1 | tail -3
timeout 300 python scripts/vasnet_perf.py 2>&1 | head -2
python bench.py --steps 3 --warmup 3 --videos 2000 --cpu-seconds 2 > gpurun_out/r1e_bench.json 2> gpurun_out/r1e_bench.err; tail -3 gpurun_out/r1e_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r1e_bench.json')); print(d['value'], d['stages_ms'], d['roofline']['frac'], d['e2e'])"

set -x
python bench.py --steps 3 --warmup 3 --videos 2000 --cpu-seconds 2 > gpurun_out/r1c_bench.json 2> gpurun_out/r1c_bench.err; tail -3 gpurun_out/r1c_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r1c_bench.json')); print(d['value'], d['stages_ms'], d['roofline']['frac'], d['train'])"

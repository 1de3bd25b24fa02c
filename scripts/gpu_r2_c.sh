#!/bin/bash
# round 2, call C: fused evaluation tail — tests, timing, bench
set -x
mkdir -p gpurun_out
python -m pytest tests/test_eval_gpu.py tests/test_trainer_gpu.py -x -q -m gpu 2>&1 | tail -5
python scripts/eval_perf.py 10000 2>&1 | tail -1 > gpurun_out/r2c_eval_perf.jsonl
SCORES=sigmoid python scripts/eval_perf.py 10000 2>&1 | tail -1 >> gpurun_out/r2c_eval_perf.jsonl
SMZ_NO_FUSED_TAIL=1 python scripts/eval_perf.py 10000 2>&1 | tail -1 >> gpurun_out/r2c_eval_perf.jsonl
cat gpurun_out/r2c_eval_perf.jsonl
python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench_n1.json 2>gpurun_out/r2c_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_n1.json')); print(d['value'], d['stages_ms'], d['roofline_eval'], d['e2e']['value'])"
tail -3 gpurun_out/r2c_bench.err

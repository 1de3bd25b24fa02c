#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_eval_gpu.py tests/test_trainer_gpu.py -x -q -m gpu 2>&1 | tail -3
python scripts/eval_perf.py 10000 2>&1 | tail -1 > gpurun_out/r2h_eval_perf.jsonl
SMZ_NO_FUSED_TAIL=1 python scripts/eval_perf.py 10000 2>&1 | tail -1 >> gpurun_out/r2h_eval_perf.jsonl
cat gpurun_out/r2h_eval_perf.jsonl
python -m pytest tests/test_train_golden_gpu.py -x -q -m gpu -s 2>&1 | tail -30
python bench.py --steps 5 --warmup 3 > gpurun_out/r2h_bench_n1.json 2>gpurun_out/r2h_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2h_bench_n1.json')); print(d['value'], d['stages_ms'], d['roofline_eval'], d['e2e']['value'], d['clocks'])"
tail -3 gpurun_out/r2h_bench.err

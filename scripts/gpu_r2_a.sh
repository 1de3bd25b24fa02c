#!/bin/bash
# round 2, call A: new eval tests, eval-path timing (dp16 on/off, slice counts), short bench
set -x
mkdir -p gpurun_out
python -m pytest tests/test_eval_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2a_pytest_eval.txt
cat gpurun_out/r2a_pytest_eval.txt
python scripts/eval_perf.py 10000 > gpurun_out/r2a_eval_perf.jsonl 2>gpurun_out/r2a_eval_perf.err
SMZ_NO_DP16=1 python scripts/eval_perf.py 10000 >> gpurun_out/r2a_eval_perf.jsonl 2>>gpurun_out/r2a_eval_perf.err
SCORES=sigmoid python scripts/eval_perf.py 10000 >> gpurun_out/r2a_eval_perf.jsonl 2>>gpurun_out/r2a_eval_perf.err
cat gpurun_out/r2a_eval_perf.jsonl; tail -5 gpurun_out/r2a_eval_perf.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_n1.json 2>gpurun_out/r2a_bench.err
cat gpurun_out/r2a_bench_n1.json; tail -5 gpurun_out/r2a_bench.err

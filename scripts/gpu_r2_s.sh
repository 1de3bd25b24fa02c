#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/r2s_bench_n1.json 2>gpurun_out/r2s_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r2s_bench_n1.json')); print(d['value'], d['e2e'], d['train'], d.get('cv_fold_parallel'), d['cpu_baseline']['stock_torch_on_this_gpu'])"
tail -3 gpurun_out/r2s_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2s_bench_ref_n1.json 2>gpurun_out/r2s_bench_ref.err
cut -c1-1500 gpurun_out/r2s_bench_ref_n1.json; tail -2 gpurun_out/r2s_bench_ref.err

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -s 2>&1 | grep -E "^E|FAILED|passed|failed|rel err" | head -30 | tee gpurun_out/r2r_pytest.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/r2r_bench_n1.json 2>gpurun_out/r2r_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2r_bench_n1.json')); print(d['value'], d['stages_ms'], d['roofline']['frac'], d['roofline']['frac_executed'], d['roofline_eval']['eval_path_frac'], d['e2e']['value'], d['clocks'], d['gpu_launches'], d['train'])"
tail -3 gpurun_out/r2r_bench.err

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2m_pytest.txt
cat gpurun_out/r2m_pytest.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/r2m_bench_n1.json 2>gpurun_out/r2m_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2m_bench_n1.json')); print(d['value'], d['stages_ms'], d['roofline'], d['roofline_eval'], d['e2e']['value'], d['clocks'])"
tail -3 gpurun_out/r2m_bench.err

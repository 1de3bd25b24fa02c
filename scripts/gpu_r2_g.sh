#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_eval_gpu.py -x -q -m gpu 2>&1 | tail -3
python scripts/eval_perf.py 10000 2>&1 | tail -1 > gpurun_out/r2g_eval_perf.jsonl
SMZ_NO_FUSED_TAIL=1 python scripts/eval_perf.py 10000 2>&1 | tail -1 >> gpurun_out/r2g_eval_perf.jsonl
cat gpurun_out/r2g_eval_perf.jsonl
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"dp16_kernel|fscore_kernel" -s 9 -c 3 --csv --log-file gpurun_out/r2g_inst.csv python scripts/eval_perf.py 10000 > /dev/null 2>&1
python - <<'PY'
import csv
for f in ('gpurun_out/r2g_inst.csv',):
    rows = list(csv.reader(open(f)))
    h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    H = rows[h]
    for r in rows[h+1:]:
        if len(r) == len(H):
            print(r[0], r[H.index('Kernel Name')][:36], r[H.index('Metric Name')], r[H.index('Metric Value')])
PY
python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_bench_n1.json 2>gpurun_out/r2g_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2g_bench_n1.json')); print(d['value'], d['stages_ms'], d['roofline_eval'], d['e2e']['value'], d['clocks'])"
tail -3 gpurun_out/r2g_bench.err

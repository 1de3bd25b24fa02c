#!/bin/bash
# round 2, call E: leaner DP loop + F-score streaming — tests, timing, instruction counts
set -x
mkdir -p gpurun_out
python -m pytest tests/test_eval_gpu.py tests/test_trainer_gpu.py -x -q -m gpu 2>&1 | tail -5
python scripts/eval_perf.py 10000 2>&1 | tail -1 > gpurun_out/r2e_eval_perf.jsonl
SMZ_NO_FUSED_TAIL=1 python scripts/eval_perf.py 10000 2>&1 | tail -1 >> gpurun_out/r2e_eval_perf.jsonl
cat gpurun_out/r2e_eval_perf.jsonl
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"dp16_kernel|fscore_kernel|pool_smem" -s 6 -c 8 --csv --log-file gpurun_out/r2e_inst.csv python scripts/eval_perf.py 10000 > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r2e_inst.csv')))
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[h]
for r in rows[h+1:]:
    if len(r) == len(H):
        print(r[H.index('Kernel Name')][:40], r[H.index('Metric Name')], r[H.index('Metric Value')])
PY

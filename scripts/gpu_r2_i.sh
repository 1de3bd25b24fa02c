#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_train_golden_gpu.py tests/test_sweep_golden_gpu.py -q -m gpu -s 2>&1 | grep -v "^$" | tail -60
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"pool_|dp16|order_|summary|fscore" -s 20 -c 6 --csv --log-file gpurun_out/r2i_inst.csv python scripts/eval_perf.py 10000 > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r2i_inst.csv')))
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[h]
for r in rows[h+1:]:
    if len(r) == len(H):
        print(r[0], r[H.index('Kernel Name')][:36], r[H.index('Metric Name')], r[H.index('Metric Value')])
PY

#!/bin/bash
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed --clock-control none -k regex:"dp16_kernel" -s 9 -c 2 --csv --log-file gpurun_out/r2f_inst.csv python scripts/eval_perf.py 10000 > /dev/null 2>&1
SMZ_NO_FUSED_TAIL=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"dp16_kernel|summary_kernel|fscore_kernel" -s 9 -c 4 --csv --log-file gpurun_out/r2f_inst_unfused.csv python scripts/eval_perf.py 10000 > /dev/null 2>&1
python - <<'PY'
import csv
for f in ('gpurun_out/r2f_inst.csv', 'gpurun_out/r2f_inst_unfused.csv'):
    rows = list(csv.reader(open(f)))
    h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    H = rows[h]
    for r in rows[h+1:]:
        if len(r) == len(H):
            print(r[0], r[H.index('Kernel Name')][:36], r[H.index('Metric Name')], r[H.index('Metric Value')])
PY

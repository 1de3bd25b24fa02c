#!/bin/bash
# round 2, call B: eval tests, eval-path timing with the resident F-score grid, per-kernel times (ncu), bench
set -x
mkdir -p gpurun_out
python -m pytest tests/test_eval_gpu.py -x -q -m gpu 2>&1 | tail -5
for c in 2 3 4; do
  SMZ_FSCORE_CTAS=$c python scripts/eval_perf.py 10000 2>&1 | tail -1 | sed "s/^/ctas=$c /" >> gpurun_out/r2b_eval_perf.jsonl
done
cat gpurun_out/r2b_eval_perf.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_eval.csv python scripts/eval_perf.py 10000 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2b_launches_eval.csv')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[hdr + 2:]:
    if len(r) > vi:
        agg[r[ki][:60]].append(float(r[vi].replace(',', '')))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    v2 = sorted(v)
    print(f"{k:60s} n={len(v):4d} median={v2[len(v2)//2]/1e3:9.1f} us max={v2[-1]/1e3:9.1f} us")
PY
python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench_n1.json 2>gpurun_out/r2b_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2b_bench_n1.json')); print(d['value'], d['stages_ms'], d['roofline_eval'], d['e2e']['value'])"
tail -3 gpurun_out/r2b_bench.err

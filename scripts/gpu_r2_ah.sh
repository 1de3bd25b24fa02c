#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
  for full in "" 1; do
    if [ -n "$full" ]; then export SMZ_EVAL_FULL_GRID=1; else unset SMZ_EVAL_FULL_GRID; fi
    python scripts/eval_perf.py 10000 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('full_grid=${full:-0}', {k: round(v,4) for k,v in d.items() if k in ('select_ms','evaluate_ms','eval_path_frac_of_6541.8GBs')})"
  done
done
unset SMZ_EVAL_FULL_GRID
for n in 9000 12000; do python scripts/eval_perf.py $n 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print($n, {k: round(v,4) for k,v in d.items() if k in ('select_ms','evaluate_ms','eval_path_frac_of_6541.8GBs')})"; SMZ_EVAL_FULL_GRID=1 python scripts/eval_perf.py $n 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print($n, 'full', {k: round(v,4) for k,v in d.items() if k in ('select_ms','evaluate_ms','eval_path_frac_of_6541.8GBs')})"; done

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r2ab_pytest.txt
tail -6 gpurun_out/r2ab_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ab_smoke.txt 2>&1; tail -3 gpurun_out/r2ab_smoke.txt
timeout 1500 python bench.py > gpurun_out/r2ab_bench_n1.json 2> gpurun_out/r2ab_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2ab_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['stages_ms'], d['roofline']['frac'], d['roofline_eval'].get('eval_path_frac'), d['e2e']['value'], d['clocks']); print(d['train']); print(d.get('cv_fold_parallel'))"
tail -3 gpurun_out/r2ab_bench.err

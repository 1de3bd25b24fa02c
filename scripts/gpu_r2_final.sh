#!/bin/bash
# final state of round 2: tests, smoke, both bench arms on one B200
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02i_pytest_gpu.txt; tail -3 gpurun_out/r02i_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02i_smoke.txt 2>&1; tail -4 gpurun_out/r02i_smoke.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02i_bench_reference_n1.json 2> gpurun_out/r02i_bench_ref.err; tail -c 600 gpurun_out/r02i_bench_reference_n1.json; echo
timeout 1500 python bench.py > gpurun_out/r02i_bench_n1.json 2> gpurun_out/r02i_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r02i_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['stages_ms'], d['roofline']['frac'], d['roofline_eval'].get('eval_path_frac'), d['e2e']['value'], d['clocks']); print({k: v for k, v in d['train'].items() if 'frames_per_s' in k or 'concurrent' in k}); print(d.get('cv_fold_parallel'))"
tail -3 gpurun_out/r02i_bench.err

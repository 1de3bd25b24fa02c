#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2v_pytest.txt
tail -8 gpurun_out/r2v_pytest.txt
for pdl in 1 0; do
  SMZ_PDL=$pdl timeout 600 python -c "
import bench, torch, json
print(json.dumps(bench.train_stage(torch.device('cuda'))))" > gpurun_out/r2v_train_pdl$pdl.json 2> gpurun_out/r2v_train_pdl$pdl.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2v_train_pdl$pdl.json').read().strip().splitlines()[-1]); print('PDL=$pdl', {k: (round(v) if isinstance(v, float) else v) for k, v in d.items() if 'frames_per_s' in k and not isinstance(v, dict)})"
  tail -3 gpurun_out/r2v_train_pdl$pdl.err
done
for pdl in 1 0; do
  SMZ_PDL=$pdl timeout 900 python bench.py --videos 4000 --steps 3 --warmup 3 --cpu-seconds 2 > gpurun_out/r2v_bench_pdl$pdl.json 2> gpurun_out/r2v_bench_pdl$pdl.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2v_bench_pdl$pdl.json').read().strip().splitlines()[-1]); print('PDL=$pdl', d['value'], d['stages_ms'], d['roofline_eval'].get('eval_path_frac'), d['e2e']['value'], d['clocks'])"
  tail -3 gpurun_out/r2v_bench_pdl$pdl.err
done

#!/bin/bash
mkdir -p gpurun_out
for pair in 1 0; do
  SMZ_GEMM_PAIR=$pair timeout 600 python -c "
import bench, torch, json
print(json.dumps(bench.train_stage(torch.device('cuda'))))" > gpurun_out/r2w_train_pair$pair.json 2> gpurun_out/r2w_train_pair$pair.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2w_train_pair$pair.json').read().strip().splitlines()[-1]); print('PAIR=$pair', {k: (round(v) if isinstance(v, float) else v) for k, v in d.items() if 'frames_per_s' in k and not isinstance(v, dict)})"
  tail -3 gpurun_out/r2w_train_pair$pair.err
done

#!/bin/bash
mkdir -p gpurun_out
for k in 4 2 6; do
  SMZ_BENCH_CONCURRENT_FOLDS=$k timeout 600 python -c "
import bench, torch, json
print(json.dumps(bench.train_stage(torch.device('cuda'))))" > gpurun_out/r2y_train_k$k.json 2> gpurun_out/r2y_train_k$k.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2y_train_k$k.json').read().strip().splitlines()[-1]); print('K=$k', round(d['vasnet_train_frames_per_s']), d['vasnet_train_concurrent_folds'])"
  tail -3 gpurun_out/r2y_train_k$k.err
done

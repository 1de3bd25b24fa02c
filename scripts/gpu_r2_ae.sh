#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2ae_pytest.txt
tail -6 gpurun_out/r2ae_pytest.txt
timeout 600 python -c "
import bench, torch, json
print(json.dumps(bench.train_stage(torch.device('cuda'))))" > gpurun_out/r2ae_train.json 2> gpurun_out/r2ae_train.err
python -c "
import json; d=json.loads(open('gpurun_out/r2ae_train.json').read().strip().splitlines()[-1]); print({k: (round(v) if isinstance(v, float) else v) for k, v in d.items() if 'frames_per_s' in k and not isinstance(v, dict)}, d['vasnet_train_concurrent_folds'])"
tail -3 gpurun_out/r2ae_train.err

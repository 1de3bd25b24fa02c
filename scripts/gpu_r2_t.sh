#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29873 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2ac_bench_n$N.json 2>gpurun_out/r2ac_bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r2ac_bench_n$N.json')); print(d['value'], d['e2e']['value'], d['e2e']['frac_of_h2d_ceiling'], d.get('strong'), d.get('train_dp'), d.get('cv_fold_parallel'))"
tail -5 gpurun_out/r2ac_bench_n$N.err

#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29877 scripts/train_dp_only.py > gpurun_out/r2ak_dp_n$N.json 2> gpurun_out/r2ak_dp_n$N.err
cat gpurun_out/r2ak_dp_n$N.json | cut -c1-1500; tail -5 gpurun_out/r2ak_dp_n$N.err | cut -c1-400

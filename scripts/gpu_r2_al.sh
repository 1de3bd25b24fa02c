#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29879 scripts/dp_graph_check.py > gpurun_out/r2al_dp_n$N.json 2> gpurun_out/r2al_dp_n$N.err
cat gpurun_out/r2al_dp_n$N.json | cut -c1-1500; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2al_dp_n$N.err | tail -12 | cut -c1-300

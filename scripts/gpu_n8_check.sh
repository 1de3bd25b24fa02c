set -x
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29831 bench.py --gpus $N --steps 3 --warmup 3 --videos 2000 > gpurun_out/r1n_bench_n$N.json 2> gpurun_out/r1n_bench_n$N.err
wc -l gpurun_out/r1n_bench_n$N.json; python -c "
import json; d=json.load(open('gpurun_out/r1n_bench_n$N.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['stages_ms'], d['e2e']['value'], d['e2e']['value_float32_rows'], d['e2e']['value_host_packed'], d['clocks'])"
tail -3 gpurun_out/r1n_bench_n$N.err

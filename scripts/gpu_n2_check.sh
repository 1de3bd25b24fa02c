set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29821 bench.py --gpus 2 --steps 3 --warmup 3 --videos 2000 > gpurun_out/r1n_bench_n2.json 2> gpurun_out/r1n_bench_n2.err
wc -l gpurun_out/r1n_bench_n2.json; python -c "
import json; d=json.load(open('gpurun_out/r1n_bench_n2.json')); print(d['n_gpus'], d['value'], d['stages_ms'], d['e2e'], d['clocks'])"
tail -3 gpurun_out/r1n_bench_n2.err
cd summarizer_b200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29823 main.py -m sumgan -s splits/summe_splits_overfit.json -c yes -e 2 -t 1 --data_parallel --pretrain_vae 1 > ../gpurun_out/r1n_sumgan_dp2.log 2>&1
grep -E "Epoch|Pretrain|Fold|Error|error" ../gpurun_out/r1n_sumgan_dp2.log | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29825 main.py -m vasnet -s splits/summe_splits.json -c yes -e 4 -t 2 > ../gpurun_out/r1n_vasnet_folds2.log 2>&1
grep -E "Fold|Cross|Error|error" ../gpurun_out/r1n_vasnet_folds2.log | tail -8

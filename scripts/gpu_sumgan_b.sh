set -x
timeout 300 python -m pytest tests/test_sumgan_gpu.py -x -q 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -6
cd summarizer_b200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 main.py -m sumgan -s splits/summe_splits_overfit.json -c yes -e 2 -t 1 --data_parallel --pretrain_vae 1 > ../gpurun_out/sumgan_dp2.log 2>&1
tail -25 ../gpurun_out/sumgan_dp2.log

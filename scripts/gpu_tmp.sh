timeout 600 python -m pytest tests/test_trainer_gpu.py tests/test_train_golden_gpu.py tests/test_dsn_gpu.py tests/test_sumgan_gpu.py -x -q -m gpu 2>&1 | tail -5
python scripts/dev/cv_time.py 2>&1 | grep "^rank"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29872 scripts/dev/cv_time.py 2>&1 | grep "^rank"

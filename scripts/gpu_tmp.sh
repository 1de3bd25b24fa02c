timeout 600 python -m pytest tests/test_eval_gpu.py tests/test_trainer_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python scripts/eval_perf.py 10000 2>&1 | tail -1

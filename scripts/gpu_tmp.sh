python -m pytest tests/test_gemm_gpu.py tests/test_vasnet_gpu.py tests/test_sweep_golden_gpu.py tests/test_trainer_gpu.py -x -q -m gpu -s 2>&1 | grep -E "^E|FAILED|passed|failed|rel err" | head -30
python scripts/dev/vas_err.py 2>&1 | tail -18
python scripts/vasnet_perf.py 2>&1 | head -4
python scripts/vasnet_steps.py 2>&1 | tail -7

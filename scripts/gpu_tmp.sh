python -m pytest tests/test_vasnet_gpu.py -x -q -m gpu 2>&1 | grep -E "^E|FAILED|passed|failed" | head -30

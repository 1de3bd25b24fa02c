timeout 1200 python -m pytest tests/test_optim_gpu.py tests/test_trainer_gpu.py tests/test_train_golden_gpu.py tests/test_dsn_gpu.py tests/test_sumgan_gpu.py tests/test_vasnet_backward_gpu.py -x -q -m gpu 2>&1 | grep -E "^E   .*(assert|Error)|passed|failed" | cut -c1-300 | head -20
python - <<'PY'
import sys, json
sys.path.insert(0, '.')
import torch, bench
print(json.dumps(bench.train_stage(torch.device('cuda', 0)))[:900])
PY

set -x
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
bash scripts/gpu_round1_d.sh

timeout 600 python -m pytest tests/test_eval_gpu.py tests/test_trainer_gpu.py tests/test_abi.py -x -q -m gpu 2>&1 | tail -3
for cfg in "" "SMZ_NO_FUSED_POOL=1"; do echo "== $cfg"; env $cfg timeout 300 python scripts/eval_perf.py 10000 2>&1 | tail -1; done
for cfg in "A=1" "SMZ_NO_FUSED_POOL=1"; do
echo "== $cfg"
env $cfg python bench.py --steps 5 --warmup 3 --cpu-seconds 1 > gpurun_out/tmp_bench.json 2>gpurun_out/tmp_bench.err
python -c "
import json; d=json.load(open('gpurun_out/tmp_bench.json')); print(d['value'], d['stages_ms'], d['roofline_eval']['eval_path_frac'], d['clocks']['sm_mhz'])"
done

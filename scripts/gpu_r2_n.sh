#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gemm_gpu.py tests/test_vasnet_gpu.py tests/test_sweep_golden_gpu.py tests/test_vasnet_backward_gpu.py tests/test_trainer_gpu.py -x -q -m gpu 2>&1 | grep -E "^E|FAILED|passed|failed" | head -30
for cfg in "" "SMZ_VASNET_LOGIT_MELEMS=68"; do
  echo "== $cfg"
  env $cfg python scripts/vasnet_perf.py 2>&1 | head -2
done 2>&1 | tee gpurun_out/r2n_vasnet_variants.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/r2n_bench_n1.json 2>gpurun_out/r2n_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_n1.json')); print(d['value'], d['stages_ms'], d['roofline']['frac'], d['roofline_eval']['eval_path_frac'], d['e2e']['value'], d['clocks'])"
tail -3 gpurun_out/r2n_bench.err
SMZ_VASNET_LOGIT_MELEMS=68 python bench.py --steps 5 --warmup 3 > gpurun_out/r2n_bench_n1_68.json 2>gpurun_out/r2n_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_n1_68.json')); print(d['value'], d['stages_ms'], d['roofline']['frac'], d['roofline_eval']['eval_path_frac'], d['e2e']['value'], d['clocks'])"

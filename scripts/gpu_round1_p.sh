set -x
P=${1:-r1p}
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > gpurun_out/${P}_pytest.log; cat gpurun_out/${P}_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; tail -2 gpurun_out/${P}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${P}_bench.json')); print(d['value'], d['stages_ms'], d['roofline']['frac'], d['roofline_eval']['eval_path_frac'], d['e2e']['value'], d['train'], d['cpu_baseline']['value'], d['clocks'])"

set -x
P=${1:-r1o}
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > gpurun_out/${P}_pytest.log; cat gpurun_out/${P}_pytest.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${P}_bench_ref.json 2> gpurun_out/${P}_bench_ref.err
python bench.py --steps 5 --warmup 3 > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; tail -3 gpurun_out/${P}_bench.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 > gpurun_out/${P}_smoke.log; cat gpurun_out/${P}_smoke.log
ncu --set full --clock-control none --import-source on -k regex:"lstm_seq_fwd_kernel|lstm_seq_bwd_kernel|lstm_decode_fwd_kernel|lstm_decode_bwd_kernel" -s 40 -c 10 -o gpurun_out/${P}_lstm_full -f python scripts/sumgan_perf.py > gpurun_out/${P}_ncu_full3.log 2>&1
python -c "
import json; d=json.load(open('gpurun_out/${P}_bench.json')); print(d['value'], d['frames_per_s'], d['stages_ms'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline_eval']['eval_path_frac'], d['e2e']['value'], d['train'], d['cpu_baseline']['value'], d['cpu_baseline']['stock_torch_on_this_gpu'], d['clocks'], d['gpu_launches'])"
cat gpurun_out/${P}_bench_ref.json | cut -c1-200

echo "== old"; LD_LIBRARY_PATH=scripts/dev/old scripts/dev/gemm_epi_bench 2>&1 | tail -20 > /dev/null
echo "== new"; LD_LIBRARY_PATH=summarizer_b200 scripts/dev/gemm_epi_bench 2>&1 | tail -20
echo "== old"; LD_LIBRARY_PATH=scripts/dev/old scripts/dev/gemm_epi_bench 2>&1 | tail -20 | grep -E "proj|pv'|k1    head \+ LN_FOLD \(|EXP|qkv"
python -m pytest tests/test_gemm_gpu.py tests/test_vasnet_gpu.py -x -q -m gpu 2>&1 | tail -2
python scripts/vasnet_perf.py 2>&1 | head -4

#!/bin/bash
mkdir -p gpurun_out
for cfg in "" "SMZ_VASNET_NO_FALLBACK=1" "SMZ_VASNET_LOGIT_MELEMS=68" "SMZ_VASNET_LOGIT_MELEMS=38" "SMZ_VASNET_LOGIT_MELEMS=68 SMZ_VASNET_NO_FALLBACK=1"; do
  echo "== $cfg"
  env $cfg python scripts/vasnet_perf.py 2>&1 | head -2
  env $cfg python scripts/vasnet_steps.py 2>&1 | tail -7
done 2>&1 | tee gpurun_out/r2k_vasnet_variants.txt

"""summarizer_b200 — B200-native (sm_100a) implementation of the hot path of sylvainma/Summarizer:
frame-importance scoring over 1024-d features -> shot selection (segment pooling + 0/1 knapsack)
-> per-user F-score, behind the reference's Python surface.

Module map (mirrors /root/reference/summarizer/):
  summarizer_b200.utils.eval      upsample, generate_scores, evaluate_scores, generate_summary, evaluate_summary
  summarizer_b200.utils.knapsack  knapsack_ortools
  summarizer_b200.models.*        VASNet, DSN, ... with forward((T,B,1024)) -> (T,B,1)
  summarizer_b200.batch           VideoBatch: device-resident ragged batches (the batched entry points)
The `summarizer` alias package at the repo root re-exports these under the reference's module paths.
All compute goes through libsummarizer_b200.so (include/summarizer_b200.h); there is no CPU fallback.
"""
__version__ = "0.1"

#!/bin/bash
# attention backward: parity + timing, persistent grid (default) vs one CTA per work item (VITK_ATTN_BWD_PERSIST=0)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_attention.py -q -m gpu -x --timeout 300 2>&1 | grep -E "passed|failed|^E |Error" | tail -8
VITK_ATTN_BWD_PERSIST=0 timeout 300 python -m pytest tests/test_gpu_attention.py -q -m gpu -x --timeout 300 2>&1 | grep -E "passed|failed|^E |Error" | tail -8
for i in 1 2; do
  timeout 60 python tools/bench_attn.py
  VITK_ATTN_BWD_PERSIST=0 timeout 60 python tools/bench_attn.py | sed 's/^/one CTA per item: /'
done 2>&1 | grep -v Warning | tee gpurun_out/bench_attn_persist.txt

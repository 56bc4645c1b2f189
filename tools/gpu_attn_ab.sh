#!/bin/bash
# attention backward: parity + timing (new attn_delta16 kernel vs the legacy one-warp-per-row kernel)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_attention.py -q -m gpu -x --timeout 300 2>&1 | grep -E "passed|failed|^E " | tail -8
for i in 1 2; do
  timeout 60 python tools/bench_attn.py
  VITK_ATTN_DELTA_LEGACY=1 timeout 60 python tools/bench_attn.py | sed 's/^/legacy delta: /'
done 2>&1 | grep -v Warning | tee gpurun_out/bench_attn_delta.txt

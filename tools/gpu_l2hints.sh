#!/bin/bash
# A/B of the L2 eviction hints (VITK_L2_HINTS bit mask: 1 = GELU' stored evict_first, 2 = multiplier / residual tiles
# loaded evict_first, 4 = LayerNorm-backward input loaded evict_first), alternating, REPS processes each.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_rowwise.py -q -m gpu -x --timeout 300 2>&1 | grep -E "passed|failed|^E " | tail -4
VITK_L2_HINTS=7 timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_rowwise.py -q -m gpu -x --timeout 300 2>&1 | grep -E "passed|failed|^E " | tail -4
for rep in $(seq 1 ${REPS:-3}); do
  for m in ${MASKS:-0 1 3 7}; do
    VITK_L2_HINTS=$m timeout 200 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --sustained-seconds 0 2>/dev/null > gpurun_out/bench_l2_${m}_$rep.json
    python - "$m" "$rep" gpurun_out/bench_l2_${m}_$rep.json <<'PY'
import json, sys
r = json.load(open(sys.argv[3]))
print(f"L2_HINTS={sys.argv[1]} rep {sys.argv[2]}: {r['value']:.1f} images/s  {r['ms_per_step']:.4f} ms  e2e {r['e2e']['value']:.1f}  sm {r['clocks'].get('sm_mhz')} MHz  gemm chain {r['roofline']['gemm_ms_per_step']:.3f} ms")
PY
  done
done | tee gpurun_out/bench_l2_ab.txt

#!/bin/bash
# A/B of the 16-epilogue-warp heavy-epilogue GEMM instantiations: parity (bit-equal to the 8-warp kernel), per-shape
# timing, and the training step with the instantiations switched on one family at a time (VITK_GEMM_EW16 bit mask:
# 1 = fc1 + GELU, 2 = fc2 data gradient × GELU', 4 = out-proj / fc2 + fp32 residual).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x --timeout 300 2>&1 | grep -E "passed|failed|^E " | tail -8
timeout 120 python tools/bench_gemm.py --ew 2>&1 | tee gpurun_out/bench_gemm_ew16.txt
for m in ${EW_MASKS:-0 7 1 2 4}; do
  echo "== VITK_GEMM_EW16=$m"
  VITK_GEMM_EW16=$m timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 0 2>/dev/null | tee gpurun_out/bench_ew16_$m.json | cut -c1-120
done
timeout 120 python tools/bench_attn.py --lib 2>&1 | grep -v Warning | tee gpurun_out/bench_attn_lib.txt

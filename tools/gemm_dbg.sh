#!/bin/bash
for d in 7 15 23 31; do echo "=== VITK_GEMM_DBG=$d"; VITK_GEMM_DBG=$d python tools/bench_gemm.py --quick 2>&1 | tail -14 | cut -c1-62; done

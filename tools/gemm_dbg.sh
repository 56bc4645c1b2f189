#!/bin/bash
for d in 0 1 2 3 4 7; do echo "=== VITK_GEMM_DBG=$d"; VITK_GEMM_DBG=$d python tools/bench_gemm.py --quick 2>&1 | tail -14; done

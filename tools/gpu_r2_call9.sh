#!/bin/bash
mkdir -p gpurun_out
for M in 4616 4608; do
echo "=== ViT-L shapes, M=$M (variant 0 = auto / pair, variant 1 = single-CTA)"
VITK_BG_M=$M VITK_BG_D=1024 VITK_BG_F=4096 python tools/bench_gemm.py --auto 2>&1 | tail -15 | cut -c1-90
done
echo "=== ViT-L cuBLAS"
VITK_BG_M=4616 VITK_BG_D=1024 VITK_BG_F=4096 python tools/bench_gemm.py --cublas 2>&1 | tail -15 | cut -c1-90

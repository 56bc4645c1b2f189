mkdir -p gpurun_out
timeout 120 python tools/bench_gemm.py --cublas 2>&1 | grep -v Warning | tee gpurun_out/bench_gemm_cublas_final.txt
timeout 120 python tools/bench_attn.py --lib 2>&1 | grep -v Warning | tee gpurun_out/bench_attn_lib_final.txt

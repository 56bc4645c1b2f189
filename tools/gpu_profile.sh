#!/bin/bash
# plain run first (must exit 0), then: launch list of one step, full captures of the top kernels
mkdir -p gpurun_out
python tools/profile_step.py > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
# forward GEMMs of layer 0 + first backward GEMMs (launch order: patch, then per layer qkv,out,fc1,fc2; backward starts at 49)
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm2_bf16 -s 0 -c 4 -f -o gpurun_out/prof_gemm_fwd python tools/profile_step.py > gpurun_out/ncu_gemm_fwd.log 2>&1
echo "gemm fwd rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm2_bf16 -s 48 -c 8 -f -o gpurun_out/prof_gemm_bwd python tools/profile_step.py > gpurun_out/ncu_gemm_bwd.log 2>&1
echo "gemm bwd rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_|layernorm|colsum|adamw|patchify" -c 12 -f -o gpurun_out/prof_other python tools/profile_step.py > gpurun_out/ncu_other.log 2>&1
echo "other rc=$?"
ls -la gpurun_out | head -30

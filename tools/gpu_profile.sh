#!/bin/bash
# Profiles of one training step (ViT-B/16@384, batch 16; tools/profile_step.py launches kernel by kernel):
#   1. plain run (must exit 0)            2. launch list with device times (every launch of the step)
#   3. ncu --set full of ONE dense middle layer, forward and backward, kernel family by kernel family, plus the optimizer
# Usage (under gpurun): bash tools/gpu_profile.sh    → gpurun_out/{launches.csv, prof_*.ncu-rep}
mkdir -p gpurun_out
P="python tools/profile_step.py"
$P > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv $P > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
full() {   # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/prof_$1 $P > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
# gemm2 launches: 4 per forward layer (qkv, out, fc1, fc2) → layer 5 = 20..23; backward: 48 forward + 8 (CLS-row top layer) + 8 per dense
# layer (fc2 wgrad, fc2 dgrad, fc1 wgrad, fc1 dgrad, out wgrad, out dgrad, qkv wgrad, qkv dgrad) → 4th dense layer = 80..87
full gemm_fwd "gemm2_bf16" 20 4
full gemm_bwd "gemm2_bf16" 80 8
full attn "attn_(fwd|bwd|delta|dq_store)" 20 4      # forward launches 0..11, then 4 per backward layer: skip 12 fwd + 2 layers
full ln "layernorm_(fwd|bwd)_kernel" 30 2           # 24 forward launches, then backward
full ln_fwd "layernorm_fwd" 10 1
full misc "colsum|sumsq|adamw|embed_bwd|patchify|head_" 8 12
ls -la gpurun_out/*.ncu-rep

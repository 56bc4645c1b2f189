#!/bin/bash
mkdir -p gpurun_out
python tools/profile_attn.py > gpurun_out/plain_attn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"attn_(fwd|bwd)_kernel" -s 2 -c 2 -f -o gpurun_out/prof_attn_r02 python tools/profile_attn.py > gpurun_out/ncu_attn_r02.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_attn_r02.log; ls -la gpurun_out/prof_attn_r02.ncu-rep

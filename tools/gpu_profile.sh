#!/bin/bash
# plain run first (must exit 0), then: launch list of one step, full captures of the top kernels
mkdir -p gpurun_out
python tools/profile_step.py > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -i 0 > gpurun_out/smi_query.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_bf16 -s 3 -c 3 -f -o gpurun_out/prof_gemm_fwd python tools/profile_step.py > gpurun_out/ncu_gemm_fwd.log 2>&1
echo "gemm fwd rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_bf16 -s 49 -c 4 -f -o gpurun_out/prof_gemm_bwd python tools/profile_step.py > gpurun_out/ncu_gemm_bwd.log 2>&1
echo "gemm bwd rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_ -c 4 -f -o gpurun_out/prof_attn python tools/profile_step.py > gpurun_out/ncu_attn.log 2>&1
echo "attn rc=$?"
ls -la gpurun_out

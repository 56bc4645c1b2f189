#!/bin/bash
# launch list of one training step + ncu --set full of one dense layer's attention backward kernels (final build)
mkdir -p gpurun_out
O=gpurun_out
P="python tools/profile_step.py"
timeout 120 $P > $O/plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain.log; exit 1; }
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv $P > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_(fwd|bwd|delta|dq_store)" -s 20 -c 4 -f -o $O/prof_attn $P > $O/ncu_attn.log 2>&1; echo "attn rc=$?"

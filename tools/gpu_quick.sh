#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_model.py -q -m gpu --timeout 600 2>&1 | grep -E "passed|failed|^E " | tail -5
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 0 2>/dev/null | cut -c1-200
python tools/profile_step.py > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"; grep -E "attn_cls" gpurun_out/launches.csv | cut -d, -f5,15- | head -4

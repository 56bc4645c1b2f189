#!/bin/bash
# Runs each GPU test file in its own process (a device-side trap poisons only that process).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in "$@"; do
  n=$(basename $f .py)
  timeout 600 python -m pytest $f -q -m gpu --timeout 120 -rf > gpurun_out/$n.log 2>&1
  r=$?
  echo "== $n rc=$r"; tail -n 40 gpurun_out/$n.log
  [ $r -ne 0 ] && rc=$r
done
exit $rc

#!/bin/bash
# model parity tests, smoke, then the bench (plain) — one gpurun call
mkdir -p gpurun_out
./tools/gpu_tests.sh tests/test_gpu_model.py
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json

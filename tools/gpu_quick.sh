#!/bin/bash
# quick re-validation after a kernel change: whole GPU test-suite, smoke, default bench
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 > $O/pytest_gpu_all.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 $O/pytest_gpu_all.log | head -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 400 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-200 $O/bench.json

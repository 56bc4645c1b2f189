#!/bin/bash
# what the driver runs at round end, plus the diagnostics kept under profiles/: full GPU test-suite, smoke, benches
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 > $O/pytest_gpu_all.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 $O/pytest_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-260 $O/bench.json
VITK_BENCH_SEGMENTED=1 python bench.py --no-cpu-baseline --sustained-seconds 0 2>/dev/null | cut -c1-200
python bench.py --graph --no-cpu-baseline --sustained-seconds 0 2>/dev/null | cut -c1-200
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "reference rc=$?"; cut -c1-300 $O/bench_ref.json

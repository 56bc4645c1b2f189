#!/bin/bash
# what the driver runs at round end, plus the diagnostics kept under profiles/: full GPU test-suite, smoke, benches, launch list
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 > $O/pytest_gpu_all.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 $O/pytest_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-260 $O/bench.json
python bench.py --config vitl384 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_vitl.json 2> $O/bench_vitl.err; echo "vitl rc=$?"; cut -c1-200 $O/bench_vitl.json
python bench.py --config vitb224-infer --steps 20 --warmup 3 > $O/bench_infer.json 2> $O/bench_infer.err; echo "infer rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "reference rc=$?"; cut -c1-200 $O/bench_ref.json
python tools/profile_step.py > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python tools/profile_step.py > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_cls" -c 2 -f -o $O/prof_attn_cls python tools/profile_step.py > $O/ncu_attn_cls.log 2>&1; echo "attn_cls rc=$?"

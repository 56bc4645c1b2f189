#!/bin/bash
# round 2, first GPU call: full GPU test-suite (one process per file), smoke, benches of every config, host overhead, cuBLAS A/B
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
for f in tests/test_gpu_*.py tests/test_metrics.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu --timeout 600 -rf -s > $O/$n.log 2>&1
  echo "== $n rc=$?"; grep -E "passed|failed|error" $O/$n.log | tail -2; grep -E "^(vit|VitkAdamW|tiny|RESULT)" $O/$n.log | head -12
done
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err; cut -c1-1500 $O/bench.json
python bench.py --steps 20 --warmup 5 --graph --no-cpu-baseline > $O/bench_graph.json 2> $O/bench_graph.err; echo "bench graph rc=$?"; tail -3 $O/bench_graph.err; cut -c1-400 $O/bench_graph.json
python bench.py --config vitl384 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_vitl.json 2> $O/bench_vitl.err; echo "bench vitl rc=$?"; tail -3 $O/bench_vitl.err; cut -c1-400 $O/bench_vitl.json
python bench.py --config vitb224-infer --steps 20 --warmup 3 > $O/bench_infer.json 2> $O/bench_infer.err; echo "bench infer rc=$?"; tail -3 $O/bench_infer.err; cut -c1-1600 $O/bench_infer.json
python bench.py --config vitb224-infer --steps 20 --warmup 3 --graph > $O/bench_infer_graph.json 2> $O/bench_infer_graph.err; echo "bench infer graph rc=$?"; tail -3 $O/bench_infer_graph.err; cut -c1-1600 $O/bench_infer_graph.json
python tools/host_overhead.py > $O/host_overhead.txt 2>&1; echo "host rc=$?"; head -3 $O/host_overhead.txt
python tools/bench_gemm.py --cublas > $O/bench_gemm_cublas_r02.txt 2>&1; echo "gemm rc=$?"; cat $O/bench_gemm_cublas_r02.txt
python tools/bench_attn.py > $O/bench_attn_r02.txt 2>&1; cat $O/bench_attn_r02.txt

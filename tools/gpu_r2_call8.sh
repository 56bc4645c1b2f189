#!/bin/bash
# ragged-M split + bucket-end segmentation: parity tests, then benches of all three configs
mkdir -p gpurun_out
O=gpurun_out
for f in tests/test_gpu_model.py tests/test_gpu_graph.py tests/test_gpu_rowwise.py tests/test_gpu_gemm.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu --timeout 600 -rf -s > $O/$n.log 2>&1
  echo "== $n rc=$?"; grep -E "passed|failed|error" $O/$n.log | tail -2; grep -E "^(FAILED|E  )" $O/$n.log | head -10; grep -hE "vit-L b" $O/$n.log
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -2 $O/bench.err; cut -c1-330 $O/bench.json
python bench.py --config vitl384 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_vitl.json 2> $O/bench_vitl.err; echo "bench vitl rc=$?"; tail -2 $O/bench_vitl.err; cut -c1-330 $O/bench_vitl.json
VITK_SPLIT_RAGGED=0 python bench.py --config vitl384 --steps 10 --warmup 3 --no-cpu-baseline --sustained-seconds 0 > $O/bench_vitl_nosplit.json 2> $O/bench_vitl_nosplit.err; echo "bench vitl nosplit rc=$?"; cut -c1-330 $O/bench_vitl_nosplit.json
python bench.py --config vitb224-infer --steps 20 --warmup 3 > $O/bench_infer.json 2> $O/bench_infer.err; echo "bench infer rc=$?"; tail -2 $O/bench_infer.err; python -c "
import json; d=json.load(open('$O/bench_infer.json')); print([(r['batch'], round(r['images_per_s']), round(r['tensor_frac_burst'],3)) for r in d['sweep']])"
python bench.py --config vitb224-infer --steps 20 --warmup 3 --graph > $O/bench_infer_graph.json 2> $O/bench_infer_graph.err; echo "bench infer graph rc=$?"; python -c "
import json; d=json.load(open('$O/bench_infer_graph.json')); print([(r['batch'], round(r['images_per_s']), round(r['tensor_frac_burst'],3)) for r in d['sweep']])"

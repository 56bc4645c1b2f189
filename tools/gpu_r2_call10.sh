#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_gemm.py -q -m gpu --timeout 600 -rf -s > $O/test_gpu_model.log 2>&1
echo "== tests rc=$?"; grep -E "passed|failed|error" $O/test_gpu_model.log | tail -2; grep -E "^(FAILED|E  )" $O/test_gpu_model.log | head -10
python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -2 $O/bench.err; cut -c1-300 $O/bench.json
python bench.py --config vitl384 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_vitl.json 2> $O/bench_vitl.err; echo "bench vitl rc=$?"; cut -c1-300 $O/bench_vitl.json
python tools/host_overhead.py > $O/host_overhead.txt 2>&1; head -2 $O/host_overhead.txt
VITK_PLAN_GRAPHS=0 python tools/host_overhead.py 2>&1 | head -1
P="python tools/profile_step.py"
$P > $O/plain.log 2>&1 &&
for spec in "attn_fwd attn_fwd_kernel 5 1" "opt adamw|sumsq|clip_scale 0 5" "embed embed_bwd|head_|patchify|embed_cls 0 5"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$2" -s $3 -c $4 -f -o $O/prof_$1 $P > $O/ncu_$1.log 2>&1; echo "$1 rc=$?"
done

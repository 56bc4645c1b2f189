#!/bin/bash
# round 2, call 2: new attention forward (pair-of-tiles kernel): parity, timing, exp2 variants; e2e probe; fixed tests
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu --timeout 300 -rf -x > $O/test_gpu_attention.log 2>&1
rc=$?; echo "== attention rc=$rc"; grep -E "passed|failed|error|Error" $O/test_gpu_attention.log | tail -5
if [ $rc -ne 0 ]; then tail -30 $O/test_gpu_attention.log; fi
V=chest-x-ray-vit_b200/csrc/build/variants
for n in default poly8 poly12 poly16; do
  if [ $n = default ]; then unset VITK_LIB; else export VITK_LIB=$PWD/$V/libvitk_$n.so; fi
  echo "=== attn $n"; timeout 120 python tools/bench_attn.py 2>&1 | tail -1
  if [ $n != default ]; then timeout 300 python -m pytest tests/test_gpu_attention.py -q -m gpu --timeout 120 -x 2>&1 | grep -E "passed|failed" | tail -1; fi
done
unset VITK_LIB
if [ $rc -eq 0 ]; then
for f in tests/test_gpu_model.py tests/test_gpu_graph.py tests/test_gpu_custom_ops.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu --timeout 600 -rf -s > $O/$n.log 2>&1
  echo "== $n rc=$?"; grep -E "passed|failed|error" $O/$n.log | tail -2; grep -E "^(vit|VitkAdamW|tiny|RESULT|traj|graph|after)" $O/$n.log | head -14; grep -E "^(FAILED|E  )" $O/$n.log | head -10
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err; cut -c1-600 $O/bench.json
fi
python tools/e2e_probe.py > $O/e2e_probe.txt 2>&1; echo "probe rc=$?"; cat $O/e2e_probe.txt | tail -16

#!/bin/bash
# round 2, call 3: attention-forward timeline (where a pair unit's cycles go), plan-level CUDA graphs, e2e probe
mkdir -p gpurun_out
O=gpurun_out
python tools/attn_fwd_timeline.py > $O/attn_fwd_timeline_r02.txt 2>&1; echo "timeline rc=$?"; cat $O/attn_fwd_timeline_r02.txt
for f in tests/test_gpu_model.py tests/test_gpu_graph.py tests/test_gpu_custom_ops.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu --timeout 600 -rf -s > $O/$n.log 2>&1
  echo "== $n rc=$?"; grep -E "passed|failed|error" $O/$n.log | tail -2; grep -E "^(FAILED|E  )" $O/$n.log | head -10
done
python tools/e2e_probe.py > $O/e2e_probe.txt 2>&1; echo "probe rc=$?"; cat $O/e2e_probe.txt | tail -13
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err; cut -c1-1400 $O/bench.json

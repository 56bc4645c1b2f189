#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu --timeout 300 -rf -x > $O/test_gpu_attention.log 2>&1
rc=$?; echo "== attention rc=$rc"; grep -E "passed|failed|error|Error" $O/test_gpu_attention.log | tail -5
if [ $rc -ne 0 ]; then grep -E "^E  " $O/test_gpu_attention.log | head; fi
python tools/attn_fwd_timeline.py > $O/attn_fwd_timeline_r02d.txt 2>&1; echo "timeline rc=$?"; cat $O/attn_fwd_timeline_r02d.txt
timeout 120 python tools/bench_attn.py 2>&1 | tail -1
python tools/e2e_probe.py 2>&1 | tail -12

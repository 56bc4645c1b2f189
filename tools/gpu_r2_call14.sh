#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_rowwise.py -q -m gpu --timeout 300 -rf > $O/test_gpu_attention.log 2>&1
echo "== attention+rowwise rc=$?"; grep -E "passed|failed|error" $O/test_gpu_attention.log | tail -2; grep -E "^(FAILED|E  )" $O/test_gpu_attention.log | head -10
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_graph.py tests/test_gpu_custom_ops.py -q -m gpu --timeout 600 -rf -s > $O/test_gpu_model.log 2>&1
echo "== model rc=$?"; grep -E "passed|failed|error" $O/test_gpu_model.log | tail -2; grep -E "^(FAILED|E  )" $O/test_gpu_model.log | head -10; grep -hE "vitb16_384 b|vit-L b" $O/test_gpu_model.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 0 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -2 $O/bench.err; cut -c1-200 $O/bench.json
VITK_CLS_ONLY_TOP=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 0 2>/dev/null | cut -c1-200
python bench.py --config vitb224-infer --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print([(r['batch'], round(r['images_per_s'])) for r in d['sweep']])"

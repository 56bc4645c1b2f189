#!/bin/bash
# N-GPU step time vs the number of CTAs NCCL may use for the overlapped gradient all-reduce
N=${1:-4}
mkdir -p gpurun_out
run() { # name, env assignment(s), extra args
  name=$1
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 $3 > gpurun_out/nccl_$name.json 2> gpurun_out/nccl_$name.err
  echo "$name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/nccl_$name.json'));print(round(d['value']),'img/s',round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['value']))" 2>&1 | tail -1)"
}
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/nccl_n1.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/nccl_n1.json'));print('n1',round(d['value']),'img/s',round(d['ms_per_step'],3),'ms')"
run default FOO=1 ""
run cta4 NCCL_MAX_CTAS=4 ""
run cta8 NCCL_MAX_CTAS=8 ""
run cta16 NCCL_MAX_CTAS=16 ""
run cta8_b6 NCCL_MAX_CTAS=8 "--layers-per-bucket 6"
run cta8_b1 NCCL_MAX_CTAS=8 "--layers-per-bucket 1"
grep -h "NCCL INFO.*channels\|nChannels\|NVLS" gpurun_out/nccl_default.err | head -5

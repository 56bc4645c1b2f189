#!/bin/bash
# scaling experiments at N GPUs: SMs reserved for NCCL during backward × bucket granularity
N=${1:-4}
mkdir -p gpurun_out
run() { # name, env assignment, extra args
  name=$1
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 $3 > gpurun_out/scale_$name.json 2> gpurun_out/scale_$name.err
  echo "$name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/scale_$name.json'));print(round(d['value']),'img/s',round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['value']))" 2>&1 | tail -1)"
}
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n1.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/scale_n1.json'));print('n1',round(d['value']),'img/s',round(d['ms_per_step'],3),'ms')"
run r0_b3 VITK_COMM_SMS=0 "--layers-per-bucket 3"
run r8_b3 VITK_COMM_SMS=8 "--layers-per-bucket 3"
run r16_b3 VITK_COMM_SMS=16 "--layers-per-bucket 3"
run r32_b3 VITK_COMM_SMS=32 "--layers-per-bucket 3"
run r16_b1 VITK_COMM_SMS=16 "--layers-per-bucket 1"
run r16_b6 VITK_COMM_SMS=16 "--layers-per-bucket 6"

#!/bin/bash
# N-GPU step time vs the all-reduce bucket schedule (sizes in the order layers finish backward)
N=${1:-4}
mkdir -p gpurun_out
run() { # name, extra args
  name=$1
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 $2 > gpurun_out/bucket_$name.json 2> gpurun_out/bucket_$name.err
  echo "$name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bucket_$name.json'));print(round(d['value']),'img/s',round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['value']))" 2>&1 | tail -1)"
}
python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bucket_n1.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/bucket_n1.json'));print('n1',round(d['value']),'img/s',round(d['ms_per_step'],3),'ms')"
run b3 "--layers-per-bucket 3"
run taper33321 "--layers-per-bucket 3,3,3,2,1"
run taper4422 "--layers-per-bucket 4,4,2,1,1"
run taper2s "--layers-per-bucket 2,2,2,2,2,1,1"
run b3_again "--layers-per-bucket 3"

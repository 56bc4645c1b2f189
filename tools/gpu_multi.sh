#!/bin/bash
# 2-GPU validation: NCCL test + bench at N=2 (and N=1 for the scaling ratio)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/test_gpu_multi.log 2>&1; echo "multi test rc=$?"; tail -5 gpurun_out/test_gpu_multi.log
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "n1 rc=$?"; cat gpurun_out/bench_n1.json | cut -c1-260
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N rc=$?"; tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json | cut -c1-260

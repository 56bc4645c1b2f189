#!/bin/bash
# N-GPU validation: NCCL parity test (12 layers, tapered buckets) + bench at N (and N=1 for the ratio)
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x -s > gpurun_out/test_gpu_multi_n$N.log 2>&1; echo "multi test rc=$?"; grep -E "RESULT|passed|failed" gpurun_out/test_gpu_multi_n$N.log | tail -3
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "n1 rc=$?"; cut -c1-200 gpurun_out/bench_n1.json
for SYNC in peer nccl; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --sustained-seconds 0 --sync $SYNC > gpurun_out/bench_n${N}_$SYNC.json 2> gpurun_out/bench_n${N}_$SYNC.err; echo "n$N $SYNC rc=$?"; grep -vE "NCCL INFO|OMP_NUM|\*\*\*" gpurun_out/bench_n${N}_$SYNC.err | tail -5; cut -c1-200 gpurun_out/bench_n${N}_$SYNC.json
done

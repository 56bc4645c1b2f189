#!/bin/bash
# short N-GPU sanity after a kernel / bench change: data-parallel parity test (both gradient syncs) and the bench at N with
# both syncs, exactly as the driver launches it.  Usage (gpurun --gpus N): bash tools/gpu_multi_quick.sh N
mkdir -p gpurun_out
N=${1:-2}
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -x -s > $O/test_gpu_multi_n$N.log 2>&1; echo "multi test rc=$?"; grep -E "RESULT|passed|failed" $O/test_gpu_multi_n$N.log | tail -3
for SYNC in auto nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --sustained-seconds 0 --sync $SYNC > $O/bench_n${N}_$SYNC.json 2> $O/bench_n${N}_$SYNC.err; echo "n$N $SYNC rc=$?"; grep -vE "NCCL INFO|OMP_NUM|\*\*\*|NCCL version|^$" $O/bench_n${N}_$SYNC.err | tail -3; cut -c1-180 $O/bench_n${N}_$SYNC.json
done
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 0 > $O/bench_n1.json 2> $O/bench_n1.err; echo "n1 rc=$?"; cut -c1-180 $O/bench_n1.json

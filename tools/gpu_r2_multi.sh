#!/bin/bash
# N-GPU validation: data-parallel parity test (12 layers, tapered buckets, both gradient syncs), per-bucket timing of the
# peer-memory sync, bench at N with both syncs and at N=1 on the same box.  Usage (gpurun --gpus N): bash tools/gpu_r2_multi.sh N
mkdir -p gpurun_out
N=${1:-2}
O=gpurun_out
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x -s > $O/test_gpu_multi_n$N.log 2>&1; echo "multi test rc=$?"; grep -E "RESULT|passed|failed" $O/test_gpu_multi_n$N.log | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/peer_sync_probe.py 2>&1 | grep -v -E "OMP_NUM|\*\*\*\*|NCCL version|^$" | tail -10 > $O/peer_sync_probe_n$N.txt; cat $O/peer_sync_probe_n$N.txt
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 0 > $O/bench_n1.json 2> $O/bench_n1.err; echo "n1 rc=$?"; cut -c1-200 $O/bench_n1.json
for SYNC in peer nccl; do
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=COLL,TUNING timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --sustained-seconds 0 --sync $SYNC > $O/bench_n${N}_$SYNC.json 2> $O/bench_n${N}_$SYNC.err; echo "n$N $SYNC rc=$?"; grep -vE "NCCL INFO|OMP_NUM|\*\*\*|NCCL version|^$" $O/bench_n${N}_$SYNC.err | tail -5; cut -c1-200 $O/bench_n${N}_$SYNC.json
done
grep -E "NCCL INFO AllReduce: [0-9]+ Bytes" $O/bench_n${N}_nccl.err | sort | uniq -c | sort -rn | head -8 > $O/nccl_algo_n$N.txt; cat $O/nccl_algo_n$N.txt

#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $N --steps 30 --warmup 5 > $O/bench_n$N.json 2> $O/multi.err
echo "bench rc=$?"; wc -l $O/bench_n$N.json; cut -c1-330 $O/bench_n$N.json
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/bench_ref_n$N.json 2>> $O/multi.err
echo "reference rc=$?"; wc -l $O/bench_ref_n$N.json; cut -c1-200 $O/bench_ref_n$N.json
tail -3 $O/multi.err | cut -c1-300

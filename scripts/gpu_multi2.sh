#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
for name in chair_camera_supervision chair_unsupervised; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
      scripts/ddp_check.py $name > $O/ddp_$name.json 2>> $O/multi.err
  echo "ddp $name rc=$?"
  cat $O/ddp_$name.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $N --steps 30 --warmup 5 > $O/bench_n$N.json 2>> $O/multi.err
echo "bench rc=$?"; wc -l $O/bench_n$N.json; cut -c1-200 $O/bench_n$N.json
tail -5 $O/multi.err

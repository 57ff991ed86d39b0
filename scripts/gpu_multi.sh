#!/bin/bash
# Multi-GPU session (gpurun --gpus N): scaling of the projection bench (no data-path collective) and
# the data-parallel train step (one NCCL gradient all-reduce), N = number of visible GPUs.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus=$N" | tee $O/multi_status.txt
for n in 1 $N; do
  if [ "$n" = "1" ]; then
    timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_n1.json 2>> $O/multi.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 \
        bench.py --gpus $n --steps 30 --warmup 5 > $O/bench_n$n.json 2>> $O/multi.err
  fi
  echo "bench n=$n rc=$?" | tee -a $O/multi_status.txt
  tail -c 600 $O/bench_n$n.json
done
for name in chair_camera_supervision chair_unsupervised; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
      scripts/ddp_check.py $name > $O/ddp_$name.json 2>> $O/multi.err
  echo "ddp $name rc=$?" | tee -a $O/multi_status.txt
  cat $O/ddp_$name.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 5 --workload train_supervised > $O/train_sup_n$N.json 2>> $O/multi.err
echo "train_sup rc=$?" | tee -a $O/multi_status.txt
tail -c 400 $O/train_sup_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus $N --steps 20 --warmup 5 --workload train_unsupervised > $O/train_unsup_n$N.json 2>> $O/multi.err
echo "train_unsup rc=$?" | tee -a $O/multi_status.txt
tail -c 400 $O/train_unsup_n$N.json
tail -5 $O/multi.err

#!/bin/bash
# Single-GPU extras: BASELINE config 5 sweep, configs 3/4 train steps on one GPU, f-1 GPU tests.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python scripts/sweep.py > $O/sweep.log 2>&1; echo "sweep rc=$?" | tee $O/extra_status.txt
tail -13 $O/sweep.log
timeout 600 python bench.py --steps 20 --warmup 5 --workload train_supervised > $O/train_sup_n1.json 2> $O/extra.err; echo "train_sup rc=$?" | tee -a $O/extra_status.txt
cat $O/train_sup_n1.json
timeout 600 python bench.py --steps 20 --warmup 5 --workload train_unsupervised --objects-per-rank 2 > $O/train_unsup_n1.json 2>> $O/extra.err; echo "train_unsup rc=$?" | tee -a $O/extra_status.txt
cat $O/train_unsup_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>> $O/extra.err; echo "reference rc=$?" | tee -a $O/extra_status.txt
cat $O/bench_reference.json
tail -3 $O/extra.err

#!/bin/bash
# One 8-GPU session: host-fabric probe, NCCL gradient-equivalence test, bench at N=8 (e2e + train rows with the 8-rank all-reduce),
# config-5 sweep at 2 / 4 / 8 GPUs.  gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu8_session.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{ nvidia-smi topo -m; nproc; numactl -H 2>/dev/null; grep -i "cpus_allowed_list\|mems_allowed_list" /proc/self/status; free -g; } > $O/topo_n8.txt 2>&1
timeout -s KILL 200 python scripts/pcie_scaling.py 8 > $O/pcie_scaling_n8.txt 2> $O/pcie_scaling_n8.err; echo "pcie rc=$?"; cut -c1-400 $O/pcie_scaling_n8.txt
timeout -s KILL 300 python -m pytest tests/test_gpu_ddp.py -m gpu -q -x -s > $O/pytest_ddp_n8.log 2>&1; echo "ddp test rc=$?"; grep "chair\|passed\|failed" $O/pytest_ddp_n8.log | tail -4
timeout -s KILL 400 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?"; cut -c1-300 $O/bench_n8.json
for n in 2 4 8; do
  timeout -s KILL 300 $TR --nproc-per-node $n --master-port 2953$n scripts/sweep.py --gpus $n > $O/sweep_n$n.json 2> $O/sweep_n$n.err; echo "sweep n$n rc=$?"
done
echo done

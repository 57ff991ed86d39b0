cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=gpurun_out
nvidia-smi topo -m > $O/topo_n2.txt 2>&1
timeout -s KILL 400 python -m pytest tests/test_gpu_ddp.py -m gpu -q -x -s > $O/pytest_ddp_h.log 2>&1; echo "ddp test rc=$?"; grep "chair\|passed\|failed\|Error" $O/pytest_ddp_h.log | tail -8
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_h_n2.json 2> $O/bench_h_n2.err; echo "bench n2 rc=$?"; tail -5 $O/bench_h_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_h_n2.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('e2e_device_resident_inputs'))
print(json.dumps(d.get('train'),indent=1)[:3500])
PY
timeout -s KILL 200 python scripts/pcie_scaling.py 2 > $O/pcie_scaling_n2.txt 2>&1; echo "pcie rc=$?"; cat $O/pcie_scaling_n2.txt | cut -c1-600

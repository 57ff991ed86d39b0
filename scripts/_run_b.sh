cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_ddp.py -m gpu -q -s > $O/pytest_ddp_n.log 2>&1; echo "ddp test rc=$?"; grep "^chair\|passed\|failed" $O/pytest_ddp_n.log | cut -c1-1200 | tail -6
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --workload train > $O/train_n2.json 2> $O/train_n2.err; echo "train n2 rc=$?"; tail -3 $O/train_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/train_n2.json'))
for k,v in d['train'].items(): print(k, v.get('graph'), v.get('eager'), v.get('allreduce_ms'), v.get('error'))
PY

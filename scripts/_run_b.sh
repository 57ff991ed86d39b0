cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_ddp.py -m gpu -q -s > $O/pytest_ddp_o.log 2>&1; echo "ddp test rc=$?"; grep "passed\|failed" $O/pytest_ddp_o.log | tail -2
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_o_n2.json 2> $O/bench_o_n2.err; echo "bench n2 rc=$?"; cut -c1-200 $O/bench_o_n2.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_o_n2.json'))
print(d['value'], d['e2e']['value'], {k:(v.get('graph',{}).get('ms_per_step'), v.get('error')) for k,v in d['train'].items()})
PY

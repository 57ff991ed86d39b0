#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $N --steps 30 --warmup 5 --scaling strong > $O/bench_strong_n$N.json 2> $O/multi.err
echo "bench rc=$?"; wc -l $O/bench_strong_n$N.json
python - <<PY
import json
d=json.load(open("$O/bench_strong_n$N.json"))
print(d["scaling"], d["n_gpus"], d["config"]["global_batch"], "%.1f us/step %.0f proj/s e2e %.0f" % (d["ms_per_step"]*1e3, d["value"], d["e2e"]["value"]))
PY
tail -3 $O/multi.err | cut -c1-300

#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_v.json 2> $O/bench_v.err; echo "bench rc=$?"
tail -5 $O/bench_v.err
python - <<PY
import json
d=json.load(open("$O/bench_v.json"))
print("%.1f us/step  %.0f proj/s  e2e %.0f" % (d["ms_per_step"]*1e3, d["value"], d["e2e"]["value"]))
print(json.dumps(d.get("variants"), indent=1))
PY

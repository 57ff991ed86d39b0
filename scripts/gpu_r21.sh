#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
for KN in "2=2" ""; do
  DPC_KNOBS=$KN timeout -s KILL 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_t.json 2> $O/bench_t.err; echo "bench [$KN] rc=$?"
  python - <<PY
import json
d=json.load(open("$O/bench_t.json"))
print("[$KN] %.1f us/step  %.0f proj/s  busy %s" % (d["ms_per_step"]*1e3, d["value"], d.get("kernel_busy_us")))
PY
done

#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
for KN in "" "10=2" "10=1" "" "10=2"; do
  DPC_KNOBS=$KN timeout -s KILL 240 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/bench_t.json 2> $O/bench_t.err; echo "bench [$KN] rc=$?"
  tail -3 $O/bench_t.err
  python - <<PY
import json
d=json.load(open("$O/bench_t.json"))
print("[$KN] %.1f us/step (events around launch %.1f)  %.0f proj/s  e2e %.0f (%.1f us)  busy %s" % (d["ms_per_step"]*1e3, d["ms_per_step_events_around_launch"]*1e3, d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"]*1e3, d.get("kernel_busy_us")))
PY
done

#!/bin/bash
# Round-16 session: in-graph timing events.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
for FL in "" "--no-l2-flush"; do
  timeout -s KILL 240 python bench.py --steps 50 --warmup 5 --no-cpu-baseline $FL > $O/bench_t.json 2> $O/bench_t.err; echo "bench [$FL] rc=$?"
  tail -3 $O/bench_t.err
  python - <<PY
import json
d=json.load(open("$O/bench_t.json"))
print("[$FL] %.1f us/step (events around launch %.1f)  %.0f proj/s  e2e %.0f (%.1f us)  busy %s" % (d["ms_per_step"]*1e3, d["ms_per_step_events_around_launch"]*1e3, d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"]*1e3, d.get("kernel_busy_us")))
print(d["config"]["launch"])
PY
done

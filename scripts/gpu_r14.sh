#!/bin/bash
# Round-14 session: raw grid zeroed by TMA bulk stores (knob 10 = 2) + early forward splat.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest variants" | tee $O/status.txt
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "splat_variants" > $O/pytest_r14.log 2>&1; echo "pytest rc=$?" | tee -a $O/status.txt
tail -3 $O/pytest_r14.log
DPC_KNOBS="10=2" timeout -s KILL 200 python scripts/step_timeline.py > $O/timeline_10_2.log 2>&1; head -10 $O/timeline_10_2.log
for KN in "" "10=2"; do
  TAG=$(echo "d$KN" | tr '=,' '__')
  DPC_KNOBS=$KN timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench [$KN] rc=$?" | tee -a $O/status.txt
  tail -3 $O/bench_$TAG.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_$TAG.json"))
    print("knobs [$KN]: %.1f us/step  %.0f proj/s  e2e %.0f (%.1f us)  busy %s" % (d["ms_per_step"]*1e3, d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"]*1e3, d.get("kernel_busy_us")))
    print(d.get("roofline_in_step"))
except Exception as e:
    print("knobs [$KN]: failed", e)
PY
done
echo "== done" | tee -a $O/status.txt

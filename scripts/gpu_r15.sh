#!/bin/bash
# Round-15 session: splat backward co-resident with the x/y pass of the backward (knob 4).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest variants" | tee $O/status.txt
timeout -s KILL 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "splat_variants" > $O/pytest_r15.log 2>&1; echo "pytest rc=$?" | tee -a $O/status.txt
tail -3 $O/pytest_r15.log
DPC_KNOBS="4=1" timeout -s KILL 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_fixture or full_benchmark_shape or properties" > $O/pytest_r15b.log 2>&1; echo "pytest(4=1) rc=$?" | tee -a $O/status.txt
tail -3 $O/pytest_r15b.log
DPC_KNOBS="4=1" timeout -s KILL 120 python scripts/step_timeline.py > $O/timeline_4_1.log 2>&1; head -9 $O/timeline_4_1.log
for KN in "" "4=1"; do
  TAG=$(echo "d$KN" | tr '=,' '__')
  DPC_KNOBS=$KN timeout -s KILL 240 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench [$KN] rc=$?" | tee -a $O/status.txt
  tail -3 $O/bench_$TAG.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_$TAG.json"))
    print("knobs [$KN]: %.1f us/step  %.0f proj/s  e2e %.0f (%.1f us)  busy %s" % (d["ms_per_step"]*1e3, d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"]*1e3, d.get("kernel_busy_us")))
except Exception as e:
    print("knobs [$KN]: failed", e)
PY
done
echo "== done" | tee -a $O/status.txt

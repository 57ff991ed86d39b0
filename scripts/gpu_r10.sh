#!/bin/bash
# Round-10 session: host taps in the pipeline prologues, leaner e2e step; splat phase stamps.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest" | tee $O/status.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/status.txt
tail -4 $O/pytest_gpu.log
echo "== splat phases" | tee -a $O/status.txt
timeout 300 python scripts/splat_phases.py > $O/splat_phases.log 2>&1; echo "phases rc=$?" | tee -a $O/status.txt
cat $O/splat_phases.log | tail -32
echo "== timeline" | tee -a $O/status.txt
timeout 300 python scripts/step_timeline.py > $O/timeline.log 2>&1; echo "timeline rc=$?" | tee -a $O/status.txt
head -9 $O/timeline.log
for KN in "" "5=1"; do
  TAG=$(echo "d$KN" | tr '=,' '__')
  DPC_KNOBS=$KN timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench [$KN] rc=$?" | tee -a $O/status.txt
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

#!/bin/bash
# Round-11 session: points-per-thread A/B of the splat kernels (knobs 0 / 1) after the 16-byte gathers; e2e after the
# materialize-grads fix.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== quick pytest" | tee $O/status.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_fixture or full_benchmark_shape or properties" > $O/pytest_r11.log 2>&1; echo "pytest rc=$?" | tee -a $O/status.txt
tail -2 $O/pytest_r11.log
for KN in "" "0=1,1=1" "0=2,1=2" "0=1" "1=1" "1=2"; do
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
DPC_KNOBS="0=1,1=1" timeout 300 python scripts/splat_phases.py > $O/splat_phases_ppt1.log 2>&1; tail -16 $O/splat_phases_ppt1.log
echo "== done" | tee -a $O/status.txt

#!/bin/bash
# Round-9 session: in-place x/y pass (two grids per step), 16-byte gathers, folded dL/dscale (no zeroing launch),
# deeper e2e pipeline; timeline + A/B + ncu captures of the defaults.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest (variants + benchmark shape)" | tee $O/status.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "splat_variants or full_benchmark_shape or golden_fixture or properties or dropout or standalone" > $O/pytest_r9.log 2>&1; echo "pytest rc=$?" | tee -a $O/status.txt
tail -4 $O/pytest_r9.log
echo "== timeline" | tee -a $O/status.txt
timeout 300 python scripts/step_timeline.py > $O/timeline.log 2>&1; echo "timeline rc=$?" | tee -a $O/status.txt
DPC_KNOBS=15=0 timeout 300 python scripts/step_timeline.py > $O/timeline_15_0.log 2>&1
cat $O/timeline.log; tail -12 $O/timeline_15_0.log
for KN in "" "15=0" "13=1" "11=0" "14=0"; do
  TAG=$(echo "d$KN" | tr '=,' '__')
  DPC_KNOBS=$KN timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench [$KN] rc=$?" | tee -a $O/status.txt
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_$TAG.json"))
    print("knobs [$KN]: %.1f us/step  %.0f proj/s  e2e %.0f (%.1f us)  busy %s" % (d["ms_per_step"]*1e3, d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"]*1e3, d.get("kernel_busy_us")))
except Exception as e:
    print("knobs [$KN]: failed", e)
PY
done
if [ "$1" != "quick" ]; then
echo "== ncu launch list" | tee -a $O/status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "ncu-list rc=$?" | tee -a $O/status.txt
echo "== ncu full" | tee -a $O/status.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dpc_ -s 12 -c 6 -o $O/prof_full \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu-full rc=$?" | tee -a $O/status.txt
fi
echo "== done" | tee -a $O/status.txt

#!/bin/bash
# Round-8 session: zeroing kernel + early splat prologue (knob 10), red.v4 (knob 11), per-kernel timeline.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest (variants + benchmark shape)" | tee $O/status.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "splat_variants or full_benchmark_shape or golden_fixture or properties or dropout" > $O/pytest_r8.log 2>&1; echo "pytest rc=$?" | tee -a $O/status.txt
tail -4 $O/pytest_r8.log
timeout 300 python -m pytest tests/test_chamfer.py -m gpu -x -q > $O/pytest_chamfer.log 2>&1; echo "pytest chamfer rc=$?" | tee -a $O/status.txt
tail -4 $O/pytest_chamfer.log
echo "== pcie probe" | tee -a $O/status.txt
timeout 120 python scripts/pcie_probe.py > $O/pcie.log 2>&1; cat $O/pcie.log
echo "== timeline" | tee -a $O/status.txt
timeout 300 python scripts/step_timeline.py > $O/timeline.log 2>&1; echo "timeline rc=$?" | tee -a $O/status.txt
cat $O/timeline.log
for KN in "10=0" "10=1" "10=1,11=1" "10=0,11=1"; do
  TAG=$(echo $KN | tr '=,' '__')
  DPC_KNOBS=$KN timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench $KN rc=$?" | tee -a $O/status.txt
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_$TAG.json"))
    print("knobs $KN: %.1f us/step  %.0f proj/s  e2e %.0f  stages %s" % (d["ms_per_step"]*1e3, d["value"], d["e2e"]["value"], {k: round(v*1e3,1) for k,v in d["stages_ms"].items()}))
except Exception as e:
    print("knobs $KN: failed", e)
PY
done
echo "== done" | tee -a $O/status.txt

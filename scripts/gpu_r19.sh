#!/bin/bash
# Round-19 session: fused silhouette-loss kernel in the step.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 300 python -m pytest tests/test_losses.py tests/test_gpu_parity.py tests/test_model_f1.py -m gpu -x -q -k "loss or full_benchmark_shape or golden or model or train" > $O/pytest_r19.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_r19.log
timeout -s KILL 120 python scripts/step_timeline.py > $O/timeline.log 2>&1; head -9 $O/timeline.log
timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_t.json 2> $O/bench_t.err; echo "bench rc=$?"
tail -3 $O/bench_t.err
python - <<PY
import json
d=json.load(open("$O/bench_t.json"))
print("%.1f us/step (events around launch %.1f)  %.0f proj/s  e2e %.0f (%.1f us)  busy %s" % (d["ms_per_step"]*1e3, d["ms_per_step_events_around_launch"]*1e3, d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"]*1e3, d.get("kernel_busy_us")))
print({k: round(v["ms_per_step"]*1e3,1) for k,v in d["variants"].items()})
PY

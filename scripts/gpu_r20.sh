#!/bin/bash
# Round-20 session: max-projection form of the pipelined depth-pass backward.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_families.py -m gpu -x -q > $O/pytest_r20.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_r20.log
timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_t.json 2> $O/bench_t.err; echo "bench rc=$?"
tail -3 $O/bench_t.err
python - <<PY
import json
d=json.load(open("$O/bench_t.json"))
print("%.1f us/step  %.0f proj/s  e2e %.0f (%.1f us)" % (d["ms_per_step"]*1e3, d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"]*1e3))
print({k: (round(v["ms_per_step"]*1e3,1), round(v["projections_per_s"])) for k,v in d["variants"].items()})
PY

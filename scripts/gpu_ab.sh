#!/bin/bash
# A/B of the smoothing-kernel families through bench.py (graph-captured step): level 0 = FFMA2, 2 = tcgen05 pipelines
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
for LV in 0 2; do
  DPC_TC=$LV timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_tc$LV.json 2> $O/bench_tc$LV.err; echo "bench level $LV rc=$?"
  python - <<PY
import json
d=json.load(open("$O/bench_tc$LV.json"))
print("level $LV: %.1f us/step  %.0f proj/s  e2e %.0f  stages %s launch=%s" % (d["ms_per_step"]*1e3, d["value"], d["e2e"]["value"], {k: round(v*1e3,1) for k,v in d["stages_ms"].items()}, d["config"].get("launch")))
PY
done

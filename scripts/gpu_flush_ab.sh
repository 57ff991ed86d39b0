#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
for M in write write+read; do
  timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --l2-flush-mode $M > $O/bench_$M.json 2> $O/bench_$M.err; echo "bench $M rc=$?"
  python - <<PY
import json
d=json.load(open("$O/bench_$M.json"))
print("$M: %.1f us/step  %.0f proj/s  e2e %.0f  stages %s" % (d["ms_per_step"]*1e3, d["value"], d["e2e"]["value"], {k: round(v*1e3,1) for k,v in d["stages_ms"].items()}))
PY
done
timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-l2-flush > $O/bench_nf.json 2> $O/bench_nf.err
python - <<PY
import json
d=json.load(open("$O/bench_nf.json"))
print("noflush: %.1f us/step  %.0f proj/s  stages %s" % (d["ms_per_step"]*1e3, d["value"], {k: round(v*1e3,1) for k,v in d["stages_ms"].items()}))
PY

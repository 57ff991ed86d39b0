#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
LV=${1:-2}
DPC_TC=$LV timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:dpc_tc -s 8 -c 4 -o $O/prof_tc -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_tc.log 2>&1; echo "ncu rc=$?"
tail -2 $O/ncu_tc.log | cut -c1-300

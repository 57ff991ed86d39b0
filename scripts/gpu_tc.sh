#!/bin/bash
# tensor-core kernel bring-up: diagnostics first (bounded), then parity tests + bench at the given level
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
LV=${1:-2}
DPC_TC=$LV timeout -s KILL 180 python scripts/tc_diag.py > $O/tc_diag.log 2>&1; rc=$?; echo "tc_diag level $LV rc=$rc" | tee $O/status.txt
cat $O/tc_diag.log | tail -45
if [ $rc -ne 0 ] && [ "$2" != "force" ]; then exit 0; fi
DPC_TC=$LV timeout -s KILL 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/status.txt
tail -5 $O/pytest_gpu.log
DPC_TC=$LV timeout -s KILL 600 python bench.py --steps 50 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/status.txt
cat $O/bench.json

#!/bin/bash
# tests + bench + sweep (no ncu)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee $O/status.txt
tail -4 $O/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/status.txt
cat $O/bench.json
timeout 900 python scripts/sweep.py > $O/sweep.log 2>&1; echo "sweep rc=$?" | tee -a $O/status.txt
tail -13 $O/sweep.log

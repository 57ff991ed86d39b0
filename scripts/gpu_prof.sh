#!/bin/bash
# launch list + full ncu capture of one forward+backward step (default kernel families)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > $O/ncu_bench.log 2>&1; echo "ncu-list rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:dpc_ -s 12 -c 6 -o $O/prof_full -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > $O/ncu_full.log 2>&1; echo "ncu-full rc=$?"
timeout -s KILL 300 python bench.py --steps 50 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
cat $O/bench.json | cut -c1-400

#!/bin/bash
# One parameterised GPU session (replaces the per-experiment scripts of round 1):
#   gpurun --timeout 900 -- 'bash scripts/gpu_session.sh tests bench ncufull'
# Tasks run in the order given; every task has its own timeout and writes to gpurun_out/<task>.*; status.txt has the rc's.
# Environment: TAG (suffix of the output names), PYTEST_ARGS (extra pytest selection), BENCH_ARGS, KNOBS (DPC_KNOBS).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:+_$TAG}
[ -n "$KNOBS" ] && export DPC_KNOBS="$KNOBS"
say() { echo "$@" | tee -a $O/status$T.txt; }
: > $O/status$T.txt
for task in "$@"; do
  say "== $task"
  case "$task" in
    topo)
      { nvidia-smi; nvidia-smi topo -m; nproc; lscpu | head -30; numactl -H 2>/dev/null; cat /proc/self/status | grep -i "cpus_allowed_list\|mems_allowed_list"; free -g; } > $O/topo$T.txt 2>&1 ;;
    tests)
      timeout -s KILL 1200 python -m pytest tests -m gpu -x -q $PYTEST_ARGS > $O/pytest_gpu$T.log 2>&1; say "pytest rc=$?"; tail -5 $O/pytest_gpu$T.log ;;
    headline)
      timeout -s KILL 600 python -m pytest tests/test_gpu_headline.py -m gpu -q > $O/pytest_headline$T.log 2>&1; say "headline rc=$?"; tail -5 $O/pytest_headline$T.log ;;
    smoke)
      timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke$T.log 2>&1; say "smoke rc=$?"; tail -2 $O/smoke$T.log ;;
    bench)
      timeout -s KILL 600 python bench.py --steps 50 --warmup 5 $BENCH_ARGS > $O/bench$T.json 2> $O/bench$T.err; say "bench rc=$?"; cut -c1-300 $O/bench$T.json; tail -3 $O/bench$T.err ;;
    benchq)   # quick: no CPU baseline / parity
      timeout -s KILL 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline $BENCH_ARGS > $O/bench$T.json 2> $O/bench$T.err; say "bench rc=$?"; cut -c1-300 $O/bench$T.json; tail -3 $O/bench$T.err ;;
    ref)
      timeout -s KILL 300 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_reference$T.json 2>> $O/bench$T.err; say "reference rc=$?" ;;
    timeline)
      timeout -s KILL 300 python scripts/step_timeline.py > $O/timeline$T.txt 2>&1; say "timeline rc=$?"; tail -12 $O/timeline$T.txt ;;
    nculist)
      timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches$T.csv \
        python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train > $O/ncu_list$T.log 2>&1; say "ncu-list rc=$?" ;;
    ncufull)
      timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:dpc_ -s ${NCU_SKIP:-14} -c ${NCU_COUNT:-7} -f -o $O/prof_full$T \
        python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train --no-graph > $O/ncu_full$T.log 2>&1; say "ncu-full rc=$?" ;;
    l1tex)    # request / sector counters of the two splat kernels (the request-rate bound of DESIGN 4a)
      timeout -s KILL 600 ncu --clock-control none -k regex:dpc_splat -s 6 -c 4 --csv --log-file $O/l1tex$T.csv --metrics \
gpu__time_duration.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,\
l1tex__t_requests_pipe_lsu_mem_global_op_red.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum,\
l1tex__m_xbar2l1tex_read_sectors.sum,l1tex__m_l1tex2xbar_write_sectors.sum,l1tex__t_sector_hit_rate.pct,\
lts__t_sectors_op_red.sum,lts__t_sectors_op_read.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__warps_active.avg.per_cycle_active \
        python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train --no-graph > $O/ncu_l1tex$T.log 2>&1; say "ncu-l1tex rc=$?" ;;
    memcheck|racecheck|synccheck)
      timeout -s KILL 900 compute-sanitizer --tool $task --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_$task$T.log 2>&1; say "$task rc=$?"; tail -3 $O/sanitizer_$task$T.log ;;
    racecheck_full|synccheck_full)   # one full-shape step under the tool
      timeout -s KILL 900 compute-sanitizer --tool ${task%_full} --error-exitcode 9 python scripts/one_step.py > $O/sanitizer_$task$T.log 2>&1; say "$task rc=$?"; tail -3 $O/sanitizer_$task$T.log ;;
    pcie)
      timeout -s KILL 300 python scripts/pcie_probe.py > $O/pcie$T.txt 2>&1; say "pcie rc=$?"; cat $O/pcie$T.txt ;;
    chamfer)
      timeout -s KILL 300 python scripts/chamfer_bench.py > $O/chamfer$T.json 2>&1; say "chamfer rc=$?" ;;
    sweep)
      timeout -s KILL 900 python scripts/sweep.py $SWEEP_ARGS > $O/sweep$T.json 2> $O/sweep$T.err; say "sweep rc=$?" ;;
    train)
      for w in train_supervised train_unsupervised; do
        timeout -s KILL 300 python bench.py --workload $w --steps 20 --warmup 5 > $O/${w}$T.json 2> $O/${w}$T.err; say "$w rc=$?"; cut -c1-300 $O/${w}$T.json
      done ;;
    *) say "unknown task $task" ;;
  esac
done
say "== done"

cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=gpurun_out
for k in "4=0" "4=1,17=2" "4=1,17=2,16=32" "4=1,17=0,16=4" "4=1,17=0,16=36"; do
  DPC_KNOBS=$k timeout -s KILL 60 python scripts/step_timeline.py > $O/timeline_l_$k.txt 2>&1; echo "knobs $k rc=$?"; grep -A9 "#1\|#2" $O/timeline_l_$k.txt | grep "xy_bwd\|total"
done
timeout -s KILL 200 python scripts/chamfer_bench.py > $O/chamfer_l.json 2> $O/chamfer_l.err; echo "chamfer rc=$?"; cut -c1-900 $O/chamfer_l.json
timeout -s KILL 300 ncu --clock-control none -k regex:dpc_nn_partial -c 2 --csv --log-file $O/ncu_chamfer_l.csv --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.sum,smsp__warps_active.avg.per_cycle_active,dram__bytes_read.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed python scripts/chamfer_bench.py > $O/ncu_chamfer_l.log 2>&1; echo "ncu chamfer rc=$?"

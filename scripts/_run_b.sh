cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=gpurun_out
for k in "23=500" "23=1000" "23=2000" "23=1000,14=0" "14=0"; do
  DPC_KNOBS=$k timeout -s KILL 60 python scripts/step_timeline.py > $O/timeline_s_$k.txt 2>&1; echo "knobs $k rc=$?"; grep -A8 "#1\|#2" $O/timeline_s_$k.txt | grep "splat_bwd\|total"
done

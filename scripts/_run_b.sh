cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=gpurun_out
for k in "22=1" "22=1,20=1" "22=1,19=1" "22=1,19=1,20=1" "22=0"; do
  DPC_KNOBS=$k timeout -s KILL 60 python scripts/step_timeline.py > $O/timeline_r_$k.txt 2>&1; echo "knobs $k rc=$?"; grep -A8 "#1\|#2" $O/timeline_r_$k.txt | grep "splat_bwd\|total"
done

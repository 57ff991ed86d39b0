cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=gpurun_out
for k in "18=0" "20=1" "19=1" "19=1,20=1" "18=1" "18=1,20=1" "18=1,19=1" "18=1,19=1,20=1"; do
  DPC_KNOBS=$k timeout -s KILL 60 python scripts/step_timeline.py > $O/timeline_j_$k.txt 2>&1; echo "knobs $k rc=$?"; grep -A8 "#1\|#2" $O/timeline_j_$k.txt | grep "splat_bwd\|total"
done
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_j.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_j.log

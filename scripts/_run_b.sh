cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=gpurun_out
timeout -s KILL 600 python scripts/sweep.py --cpu > $O/sweep_n1.json 2> $O/sweep_n1.err; echo "sweep rc=$?"; tail -13 $O/sweep_n1.err | cut -c1-330
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_i.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu_i.log

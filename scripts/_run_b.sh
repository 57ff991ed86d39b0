cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=gpurun_out
export DPC_KNOBS=21=1
timeout -s KILL 60 python scripts/one_step.py 2 1 > $O/one_step_m.log 2>&1; echo "one_step B=2 rc=$?"; tail -1 $O/one_step_m.log
timeout -s KILL 60 python scripts/one_step.py 32 3 >> $O/one_step_m.log 2>&1; rc=$?; echo "one_step B=32 rc=$rc"; tail -1 $O/one_step_m.log
if [ $rc -ne 0 ]; then echo "fused fwd broken: stop"; exit 0; fi
timeout -s KILL 100 python scripts/step_timeline.py > $O/timeline_m_fused.txt 2>&1; echo "timeline rc=$?"; grep -A9 "#1\|#2" $O/timeline_m_fused.txt | grep "splat_fwd\|xy_fwd\|z_fwd\|z_bwd\|total"
timeout -s KILL 200 python -m pytest tests/test_gpu_headline.py -m gpu -q -x > $O/pytest_headline_m.log 2>&1; echo "headline rc=$?"; tail -2 $O/pytest_headline_m.log
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_benchmark or clustered or max_projection or golden_fixture" > $O/pytest_parity_m.log 2>&1; echo "parity rc=$?"; tail -2 $O/pytest_parity_m.log
timeout -s KILL 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-train > $O/bench_m.json 2> $O/bench_m.err; echo "bench rc=$?"; cut -c1-260 $O/bench_m.json

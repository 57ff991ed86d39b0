cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; O=gpurun_out
timeout -s KILL 60 python scripts/one_step.py 2 1 > $O/one_step_f.log 2>&1; echo "one_step B=2 rc=$?"
timeout -s KILL 60 python scripts/one_step.py 32 3 >> $O/one_step_f.log 2>&1; rc=$?; echo "one_step B=32 rc=$rc"; tail -2 $O/one_step_f.log
if [ $rc -ne 0 ]; then echo "fused path broken: stop"; exit 0; fi
for k in "17=0" "17=1" "17=0,16=1" "17=1,16=1" "17=0,16=4" "17=1,16=4" "17=0,16=3" "17=1,16=3"; do
  DPC_KNOBS=$k timeout -s KILL 60 python scripts/step_timeline.py > $O/timeline_f_$k.txt 2>&1; echo "knobs $k rc=$?"; grep -A10 "#2" $O/timeline_f_$k.txt | grep "xy_bwd\|splat_bwd\|fused\|total"
done
timeout -s KILL 200 python -m pytest tests/test_gpu_headline.py -m gpu -q -x > $O/pytest_headline_f.log 2>&1; echo "headline rc=$?"; tail -2 $O/pytest_headline_f.log
DPC_KNOBS=17=1 timeout -s KILL 200 python -m pytest tests/test_gpu_headline.py -m gpu -q -x > $O/pytest_headline_f1.log 2>&1; echo "headline wide rc=$?"; tail -2 $O/pytest_headline_f1.log
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "splat_variants or full_benchmark or clustered or max_projection or golden_fixture" > $O/pytest_parity_f.log 2>&1; echo "parity rc=$?"; tail -2 $O/pytest_parity_f.log
timeout -s KILL 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_f.json 2> $O/bench_f.err; echo "bench rc=$?"; cut -c1-300 $O/bench_f.json
DPC_KNOBS=17=1 timeout -s KILL 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_f1.json 2> $O/bench_f1.err; echo "bench wide rc=$?"; cut -c1-300 $O/bench_f1.json

#!/bin/bash
# End-of-round session: full GPU test-suite, smoke, bench (+ reference arm), ncu launch list + full capture of one step.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi > $O/nvidia_smi.txt 2>&1; nproc > $O/nproc.txt
echo "== pytest gpu" | tee $O/status.txt
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/status.txt
tail -3 $O/pytest_gpu.log
echo "== smoke" | tee -a $O/status.txt
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/status.txt
tail -2 $O/smoke.log
echo "== bench" | tee -a $O/status.txt
timeout -s KILL 600 python bench.py --steps 50 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/status.txt
cut -c1-400 $O/bench.json
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>> $O/bench.err; echo "reference rc=$?" | tee -a $O/status.txt
echo "== ncu launch list" | tee -a $O/status.txt
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "ncu-list rc=$?" | tee -a $O/status.txt
echo "== ncu full" | tee -a $O/status.txt
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:dpc_ -s 14 -c 7 -o $O/prof_full \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu-full rc=$?" | tee -a $O/status.txt
echo "== sanitizer" | tee -a $O/status.txt
timeout -s KILL 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer.log 2>&1; echo "memcheck rc=$?" | tee -a $O/status.txt
tail -3 $O/sanitizer.log
echo "== done" | tee -a $O/status.txt

#!/bin/bash
# Round-13 session: whole-step graphs (copies inside) for e2e, backward splat at 1 point/thread, chamfer measurement.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== quick pytest" | tee $O/status.txt
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_chamfer.py -m gpu -x -q -k "golden_fixture or full_benchmark_shape or properties or chamfer or cuda_kernels" > $O/pytest_r13.log 2>&1; echo "pytest rc=$?" | tee -a $O/status.txt
tail -2 $O/pytest_r13.log
timeout -s KILL 300 python bench.py --steps 50 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/status.txt
tail -3 $O/bench.err; cat $O/bench.json
timeout -s KILL 300 python scripts/chamfer_bench.py > $O/chamfer.json 2> $O/chamfer.err; echo "chamfer rc=$?" | tee -a $O/status.txt
tail -3 $O/chamfer.err; cat $O/chamfer.json
echo "== done" | tee -a $O/status.txt

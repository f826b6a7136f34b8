#!/bin/bash
# r02 call 22: per-observation point sums in the product pass -- bitwise tests, full suite, bench
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c22; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -n 5 $O/gpu_tests.log
TIME_MODES=0,4,0 timeout 600 python tools/matvec_time.py > $O/mv_modes.log 2>&1; tail -n 3 $O/mv_modes.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.json

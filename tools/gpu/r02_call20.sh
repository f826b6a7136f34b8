#!/bin/bash
# r02 call 20: layout built on the device -- full GPU suite, e2e trace, bench
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c20; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -n 6 $O/gpu_tests.log
timeout 600 python tools/e2e_trace.py > $O/e2e_trace.log 2>&1; grep -E "preprocess|rep " $O/e2e_trace.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.json

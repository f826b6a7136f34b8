#!/bin/bash
# r02 call 18: two virtual blocks per round in the vector phases of the fused solve -- bitwise tests, N=1 bench
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c18; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "fused or bitwise" > $O/gpu_tests_fused.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests_fused.log
tail -n 6 $O/gpu_tests_fused.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.json
timeout 1200 python tools/final_scaling.py --steps 6 --warmup 2 > $O/final_n1.json 2> $O/final_n1.err; tail -c 600 $O/final_n1.json

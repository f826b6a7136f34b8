#!/bin/bash
# r02 call 27: three virtual blocks per round by lane groups -- bitwise tests, N=1 bench (no regression?), Final shape N=1
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c27; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "fused or bitwise" > $O/gpu_tests_fused.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests_fused.log
tail -n 4 $O/gpu_tests_fused.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.json
timeout 1200 python tools/final_scaling.py --steps 6 --warmup 2 > $O/final_n1.json 2> $O/final_n1.err; tail -c 300 $O/final_n1.json

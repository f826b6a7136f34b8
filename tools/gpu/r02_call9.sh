#!/bin/bash
# r02 call 9: fused PCG solve as the default -- full GPU suite, bench, ncu launch list
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c9; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -n 5 $O/gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.json; tail -n 3 $O/bench_n1.err
timeout 300 python tools/prof_one_iteration.py 7 > $O/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv python tools/prof_one_iteration.py 7 > $O/ncu_list.log 2>&1

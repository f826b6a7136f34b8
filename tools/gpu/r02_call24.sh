#!/bin/bash
# r02 call 24: Schur set-up with the 2 x 2 projector (F^T P F) -- parity suite, family times, bench
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c24; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -n 8 $O/gpu_tests.log
AB_PROFILE=1 timeout 300 python tools/matvec_ab.py > $O/ab_allfam.log 2>&1; tail -n 2 $O/ab_allfam.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.json

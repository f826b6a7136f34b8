#!/bin/bash
# r02 call 4: product with the deferred direction FMA (modes 0 / 4), serial vs chunked sums, GPU tests, bench
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c4; mkdir -p $O
TIME_MODES=0,4,0 SKERES_MATVEC_SUMS=serial timeout 600 python tools/matvec_time.py > $O/mv_modes_serial.log 2>&1; tail -n 4 $O/mv_modes_serial.log
TIME_MODES=0,4,0 SKERES_MATVEC_SUMS=chunked timeout 600 python tools/matvec_time.py > $O/mv_modes_chunked.log 2>&1; tail -n 4 $O/mv_modes_chunked.log
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -n 8 $O/gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 400 $O/bench_n1.json; tail -n 3 $O/bench_n1.err
SKERES_MATVEC_SUMS=chunked timeout 900 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/bench_n1_chunked.json 2> $O/bench_n1_chunked.err

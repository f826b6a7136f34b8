#!/bin/bash
# r02 call 3: why is the product slower inside the PCG loop than back to back?  + GPU tests + bench line
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c3; mkdir -p $O
TIME_MODES=0,1,2,3,4,5,0 SKERES_MATVEC_SUMS=serial timeout 600 python tools/matvec_time.py > $O/mv_modes_serial.log 2>&1; tail -n 8 $O/mv_modes_serial.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -n 5 $O/gpu_tests.log
SKERES_MATVEC_SUMS=serial timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1_serial.json 2> $O/bench_n1_serial.err; tail -c 600 $O/bench_n1_serial.json; tail -n 3 $O/bench_n1_serial.err

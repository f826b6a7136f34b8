#!/bin/bash
# r02 call 2: implicit Schur product variants timed in isolation + ablation floors; GPU tests
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c2; mkdir -p $O
run() { name=$1; shift; env "$@" timeout 300 python tools/matvec_time.py > $O/mv_$name.log 2>&1; tail -n 1 $O/mv_$name.log; }
run tma_chunked SKERES_MATVEC=tma
run tma_serial SKERES_MATVEC=tma SKERES_MATVEC_SUMS=serial
run rows_serial SKERES_MATVEC=rows SKERES_MATVEC_SUMS=serial
run rows_chunked SKERES_MATVEC=rows
run classic_serial SKERES_MATVEC=classic SKERES_MATVEC_SUMS=serial
for a in 1 2 3; do run ablate$a SKERES_MATVEC=tma SKERES_MATVEC_SUMS=serial SKERES_LIB=$PWD/gpurun_variants/libskeres_ablate$a.so; done
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -n 5 $O/gpu_tests.log

#!/bin/bash
# r02 call 12: FP64 tensor-core (DMMA) Cholesky update -- GPU suite, Ladybug-49 DENSE_SCHUR timing, ncu evidence of the tensor pipe
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c12; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -n 12 $O/gpu_tests.log
timeout 300 python tools/dense_schur_bench.py ladybug-49 > $O/dense_schur_ladybug.json 2> $O/dense_schur_ladybug.err; cat $O/dense_schur_ladybug.json; tail -n 3 $O/dense_schur_ladybug.err
timeout 300 python tools/dense_schur_bench.py ladybug-49 sparse > $O/sparse_schur_ladybug.json 2>> $O/dense_schur_ladybug.err; cat $O/sparse_schur_ladybug.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chol_update -s 20 -c 2 -o $O/prof_chol_update python tools/dense_schur_bench.py ladybug-49 > $O/ncu.log 2>&1
tail -3 $O/ncu.log

#!/bin/bash
# r02 call 15: where do the small dense kernels spend their time?
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c15; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_chol_panel|k_schur_offdiag|k_chol_solve" -s 30 -c 4 -o $O/prof_dense python tools/dense_schur_bench.py ladybug-49 > $O/ncu.log 2>&1
tail -3 $O/ncu.log

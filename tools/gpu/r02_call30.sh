#!/bin/bash
# r02 call 30: ncu of the set-up kernels (Jacobian evaluation, Schur set-up, back-substitution)
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c30; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ba_evaluate|k_ba_schur_setup|k_ba_back_substitute" -s 6 -c 4 -o $O/prof_setup python tools/prof_one_iteration.py 4 > $O/ncu.log 2>&1
tail -3 $O/ncu.log

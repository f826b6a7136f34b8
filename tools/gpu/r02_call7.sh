#!/bin/bash
# r02 call 7: why is the fused PCG solve slower than the kernel sequence?  variants + ncu capture
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c7; mkdir -p $O
export SKERES_PCG=fused
timeout 300 python tools/matvec_ab.py > $O/ab_fused.log 2>&1; tail -n 2 $O/ab_fused.log
for v in nocoh nochain; do
  SKERES_LIB=$PWD/gpurun_variants/libskeres_$v.so timeout 300 python tools/matvec_ab.py > $O/ab_$v.log 2>&1; tail -n 2 $O/ab_$v.log
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcg_solve -s 3 -c 1 -o $O/prof_pcg_solve python tools/prof_one_iteration.py 5 > $O/ncu.log 2>&1
tail -3 $O/ncu.log

#!/bin/bash
# r02 call 8: fused PCG solve with ONE call site of the product pass -- tests, A/B against the sequence, ncu
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c8; mkdir -p $O
SKERES_PCG=fused timeout 600 python -m pytest tests -m gpu -q -x -k "fused or bitwise" > $O/gpu_tests_fused.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests_fused.log
tail -n 5 $O/gpu_tests_fused.log
SKERES_PCG=fused timeout 300 python tools/matvec_ab.py > $O/ab_fused.log 2>&1; tail -n 2 $O/ab_fused.log
SKERES_PCG=sequence timeout 300 python tools/matvec_ab.py > $O/ab_sequence.log 2>&1; tail -n 2 $O/ab_sequence.log
SKERES_PCG=fused timeout 900 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/bench_n1_fused.json 2> $O/bench_n1_fused.err; tail -c 300 $O/bench_n1_fused.json
SKERES_PCG=fused timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcg_solve -s 3 -c 1 -o $O/prof_pcg_solve python tools/prof_one_iteration.py 5 > $O/ncu.log 2>&1
tail -3 $O/ncu.log

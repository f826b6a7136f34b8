#!/bin/bash
# r02 call 1: GPU tests, matvec A/B (serial chains vs chunked two-level sums), launch list + ncu capture of the new kernel
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r02c1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02c1/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c1/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c1/gpu_tests.log
tail -5 gpurun_out/r02c1/gpu_tests.log
for m in serial chunked; do
  SKERES_MATVEC_SUMS=$m timeout 300 python tools/matvec_ab.py > gpurun_out/r02c1/ab_$m.log 2>&1
  AB_PROFILE=1 SKERES_MATVEC_SUMS=$m timeout 300 python tools/matvec_ab.py > gpurun_out/r02c1/ab_${m}_allfam.log 2>&1
  tail -3 gpurun_out/r02c1/ab_$m.log
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c1/bench_n1.json 2> gpurun_out/r02c1/bench_n1.err; tail -c 1500 gpurun_out/r02c1/bench_n1.json
timeout 300 python tools/prof_one_iteration.py 7 > gpurun_out/r02c1/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ba_matvec_tma -s 20 -c 2 -o gpurun_out/r02c1/prof_matvec python tools/prof_one_iteration.py 7 > gpurun_out/r02c1/ncu.log 2>&1
tail -3 gpurun_out/r02c1/ncu.log

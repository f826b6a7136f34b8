#!/bin/bash
# r02 call 6: fused PCG solve (k_pcg_solve) -- bitwise tests against the kernel sequence, full GPU suite, bench A/B
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c6; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "fused or bitwise" > $O/gpu_tests_fused.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests_fused.log
tail -n 15 $O/gpu_tests_fused.log
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -n 8 $O/gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/bench_n1_fused.json 2> $O/bench_n1_fused.err; tail -c 600 $O/bench_n1_fused.json; tail -n 3 $O/bench_n1_fused.err
SKERES_PCG=sequence timeout 900 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/bench_n1_sequence.json 2> $O/bench_n1_sequence.err; tail -c 600 $O/bench_n1_sequence.json

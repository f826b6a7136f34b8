#!/bin/bash
# r02 call 32: user functors in the tile kernels (tests against the oracle's division model), batched curve fits: shipped build, more variants, ncu capture
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c32; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -k "batched or user or unseen" > $O/gpu_tests_batched_user.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests_batched_user.log
tail -n 30 $O/gpu_tests_batched_user.log
timeout 200 python tools/batch_bench.py > $O/batch_shipped.json 2> $O/batch_shipped.err; cat $O/batch_shipped.json
for v in batch_rb4_mb4 batch_rb4_mb6 batch_rb8_mb5 batch_rb3_mb5; do
  SKERES_LIB=$PWD/gpurun_variants/libskeres_$v.so timeout 200 python tools/batch_bench.py > $O/$v.json 2> $O/$v.err; cat $O/$v.json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_batch_iterate -s 2 -c 2 -o $O/prof_batch python tools/batch_bench.py > $O/ncu.log 2>&1
tail -3 $O/ncu.log

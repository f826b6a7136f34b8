#!/bin/bash
# r02 call 31: batched curve fits (configs[3]) -- two passes per LM iteration with the saved sums against the four-pass version; register-cap /
# load-batching variants; functors of the bundle-adjustment shape given as source, in the tile kernels
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c31; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "batched or user" > $O/gpu_tests_batched_user.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests_batched_user.log
tail -n 30 $O/gpu_tests_batched_user.log
timeout 200 python tools/batch_bench.py > $O/batch_new.json 2> $O/batch_new.err; cat $O/batch_new.json
for v in batch_old batch_rb1 batch_rb2_mb5 batch_rb4_mb5 batch_rb2_mb6; do
  SKERES_LIB=$PWD/gpurun_variants/libskeres_$v.so timeout 200 python tools/batch_bench.py > $O/$v.json 2> $O/$v.err; cat $O/$v.json
done
timeout 900 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -n 5 $O/gpu_tests.log

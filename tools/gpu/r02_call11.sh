#!/bin/bash
# r02 call 11: run-time compiled functors on the device + full GPU suite
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c11; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -n 30 $O/gpu_tests.log

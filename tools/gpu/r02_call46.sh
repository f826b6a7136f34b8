#!/bin/bash
# r02 call 46 (2 GPUs): multi-GPU correctness check of the final tree (all ranks bit-equal, rank-local ingestion, NaN on one rank -> FAILURE on all, multi vs single)
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c46; mkdir -p $O
export SKERES_PEER_TIMEOUT_S=15
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR tools/multi_gpu_check.py ladybug-49 > $O/check_ladybug.log 2>&1; echo "rc=$?" >> $O/check_ladybug.log; grep -E "single GPU|multi vs|MULTI|rc=|bit|NaN|FAIL" $O/check_ladybug.log | tail -12
timeout 200 $TR tools/multi_gpu_check.py wide 4 > $O/check_wide.log 2>&1; echo "rc=$?" >> $O/check_wide.log; grep -E "single GPU|multi vs|MULTI|rc=|bit|NaN|FAIL" $O/check_wide.log | tail -12

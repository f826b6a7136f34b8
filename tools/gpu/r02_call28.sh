#!/bin/bash
# r02 call 28 (2 GPUs): packed vector phases with the peer-window gather -- correctness on a problem with more virtual blocks than CTAs
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c28; mkdir -p $O
export SKERES_PEER_TIMEOUT_S=15
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR tools/multi_gpu_check.py wide 4 > $O/check_wide.log 2>&1; echo "rc=$?" >> $O/check_wide.log; grep -E "rank 0:|single GPU|multi vs|MULTI|rc=" $O/check_wide.log

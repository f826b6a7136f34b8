#!/bin/bash
# r02 call 29 (2 GPUs): own-virtual-block walk in the peer mode -- correctness (wide, ladybug), N=2 bench
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c29; mkdir -p $O
export SKERES_PEER_TIMEOUT_S=15
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR tools/multi_gpu_check.py wide 4 > $O/check_wide.log 2>&1; echo "rc=$?" >> $O/check_wide.log; grep -E "single GPU|multi vs|MULTI|rc=" $O/check_wide.log
timeout 200 $TR tools/multi_gpu_check.py ladybug-49 > $O/check_ladybug.log 2>&1; echo "rc=$?" >> $O/check_ladybug.log; grep -E "single GPU|multi vs|MULTI|rc=" $O/check_ladybug.log
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; tail -c 300 $O/bench_n2.json

#!/bin/bash
# r02 call 21 (2 GPUs): device-built layout in the rank-local mode -- correctness against one GPU / replicated ingestion, bench
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c21; mkdir -p $O
export SKERES_PEER_TIMEOUT_S=15
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/multi_gpu_check.py ladybug-49 > $O/check_ladybug.log 2>&1; echo "rc=$?" >> $O/check_ladybug.log; tail -n 6 $O/check_ladybug.log
timeout 400 $TR tools/multi_gpu_check.py venice-1778 > $O/check_venice.log 2>&1; echo "rc=$?" >> $O/check_venice.log; tail -n 6 $O/check_venice.log
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; tail -c 400 $O/bench_n2.json; tail -n 3 $O/bench_n2.err

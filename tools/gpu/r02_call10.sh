#!/bin/bash
# r02 call 10 (2 GPUs): fused PCG solve with the in-kernel peer-window exchange -- correctness against one GPU, bench A/B
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c10; mkdir -p $O
export SKERES_PEER_TIMEOUT_S=15
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/multi_gpu_check.py ladybug-49 > $O/check_ladybug.log 2>&1; echo "rc=$?" >> $O/check_ladybug.log; tail -n 12 $O/check_ladybug.log
timeout 400 $TR tools/multi_gpu_check.py venice-1778 > $O/check_venice.log 2>&1; echo "rc=$?" >> $O/check_venice.log; tail -n 12 $O/check_venice.log
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_fused.json 2> $O/bench_n2_fused.err; tail -c 400 $O/bench_n2_fused.json; tail -n 3 $O/bench_n2_fused.err
SKERES_PCG=sequence timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > $O/bench_n2_sequence.json 2> $O/bench_n2_sequence.err; tail -c 400 $O/bench_n2_sequence.json

#!/bin/bash
# r02 call 43 (N GPUs): bench.py as the driver's scaling run launches it, final tree
set -x
cd "$GRAFT_REPO_ROOT"
N=${SCALE_N:-4}   # gpurun --gpus N must match
O=gpurun_out/r02c43; mkdir -p $O
export SKERES_PEER_TIMEOUT_S=30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err
echo "rc=$?"; tail -c 400 $O/bench_n$N.json; tail -n 3 $O/bench_n$N.err

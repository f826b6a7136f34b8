#!/bin/bash
# the driver's scaling run for one N (bench.py under torchrun)
set -x
cd "$GRAFT_REPO_ROOT"
N=${SCALE_N:-8}
O=gpurun_out/r02scale; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err
echo "rc=$?"; tail -c 600 $O/bench_n$N.json; tail -n 3 $O/bench_n$N.err

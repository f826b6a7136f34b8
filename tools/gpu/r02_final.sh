#!/bin/bash
# BASELINE configs[4]: Final-13682 shape, strong scaling.  usage: r02_final.sh N
set -x
cd "$GRAFT_REPO_ROOT"
N=${FINAL_N:-8}
O=gpurun_out/r02final; mkdir -p $O
export SKERES_PEER_TIMEOUT_S=30
if [ "$N" = "1" ]; then
  timeout 1200 python tools/final_scaling.py --steps 6 --warmup 2 > $O/final_n1.json 2> $O/final_n1.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/final_scaling.py --steps 6 --warmup 2 > $O/final_n$N.json 2> $O/final_n$N.err
fi
echo "rc=$?"; tail -c 1500 $O/final_n$N.json; tail -n 5 $O/final_n$N.err

#!/bin/bash
# r02 call 40 (8 GPUs): the two multi-GPU headline runs with the final tree (L2 copy policies on by working set): Venice x 8 weak scaling, Final-13682 strong scaling
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c40; mkdir -p $O
export SKERES_PEER_TIMEOUT_S=30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29544 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "rc=$?"; tail -c 500 $O/bench_n8.json; tail -n 3 $O/bench_n8.err
timeout 900 $TR --master-port 29545 tools/final_scaling.py --steps 6 --warmup 2 > $O/final_n8.json 2> $O/final_n8.err; echo "rc=$?"; tail -c 700 $O/final_n8.json; tail -n 3 $O/final_n8.err

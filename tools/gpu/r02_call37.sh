#!/bin/bash
# r02 call 37: Final-13682 shape on one GPU with and without the L2 copy policies; Ladybug DENSE_SCHUR line; batched curve fits shipped build
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c37; mkdir -p $O
for mb in 24 -1 0 64; do
SKERES_L2_KEEP_MB=$mb timeout 600 python tools/final_scaling.py --steps 6 --warmup 2 > $O/final_n1_keep$mb.json 2> $O/final_n1_keep$mb.err; tail -c 600 $O/final_n1_keep$mb.json
done
timeout 300 python tools/dense_schur_bench.py ladybug-49 > $O/dense_ladybug.json 2> $O/dense_ladybug.err; tail -c 400 $O/dense_ladybug.json
timeout 200 python tools/batch_bench.py > $O/batch_shipped.json 2> $O/batch_shipped.err; cat $O/batch_shipped.json

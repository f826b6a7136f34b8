#!/bin/bash
# r02 call 25 (2 GPUs): where do the slow end-to-end passes at N > 1 come from?
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c25; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for i in 1 2; do
SKERES_TRACE_HOST=1 timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-profile > $O/bench_n2_$i.json 2> $O/bench_n2_$i.err
grep -E "sk_solve:|preprocess:|sk_solver_destroy|peer" $O/bench_n2_$i.err | tail -16
python -c "
import json;d=json.loads([l for l in open('$O/bench_n2_$i.json') if l.startswith('{')][-1]);print(d['e2e']['wall_s_runs'], d['e2e']['preprocessor_s'])"
done

#!/bin/bash
# r02 call 16: end-to-end host trace
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c16; mkdir -p $O
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null
for i in 1 2; do timeout 600 python tools/e2e_trace.py > $O/e2e_trace_$i.log 2>&1; grep -E "destroy|rep " $O/e2e_trace_$i.log; done
SKERES_TRACE_HOST=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; grep -E "destroy|sk_solve" $O/bench.err | tail -8

#!/bin/bash
# r02 call 42: final single-GPU validation of the round-2 tree -- smoke(), full GPU suite, bench as the driver runs it, reference arm
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c42; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -n 4 $O/gpu_tests.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.json
timeout 200 python tools/batch_bench.py > $O/batch.json 2> $O/batch.err; cat $O/batch.json

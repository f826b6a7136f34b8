#!/bin/bash
# r02 call 36: validation of the tree with the L2 copy policies as the default -- smoke(), full GPU suite, bench as the driver runs it; launch list + ncu capture of k_pcg_solve
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c36; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -n 4 $O/gpu_tests.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.json
timeout 300 python tools/prof_one_iteration.py 7 > $O/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv python tools/prof_one_iteration.py 7 > $O/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcg_solve -s 3 -c 1 -o $O/prof_pcg_solve python tools/prof_one_iteration.py 5 > $O/ncu.log 2>&1
tail -3 $O/ncu.log; cat $O/plain.log

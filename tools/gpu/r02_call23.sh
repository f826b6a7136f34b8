#!/bin/bash
# r02 call 23: ncu capture of the final k_pcg_solve (traffic per product for bench.py), launch list
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c23; mkdir -p $O
timeout 300 python tools/prof_one_iteration.py 7 > $O/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv python tools/prof_one_iteration.py 7 > $O/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcg_solve -s 3 -c 1 -o $O/prof_pcg_solve python tools/prof_one_iteration.py 5 > $O/ncu.log 2>&1
tail -3 $O/ncu.log; cat $O/plain.log

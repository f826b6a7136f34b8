#!/bin/bash
# r02 call 5: row-by-row diagnostics of the failing parity tests; launch list and ncu capture of the current product kernel
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c5; mkdir -p $O
timeout 600 python tools/long_track_rows.py full > $O/rows_full_tma.log 2>&1; tail -n 60 $O/rows_full_tma.log
SKERES_MATVEC=classic timeout 600 python tools/long_track_rows.py full > $O/rows_full_classic.log 2>&1
timeout 300 python tools/prof_one_iteration.py 7 > $O/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv python tools/prof_one_iteration.py 7 > $O/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ba_matvec_tma -s 20 -c 2 -o $O/prof_matvec python tools/prof_one_iteration.py 7 > $O/ncu.log 2>&1
tail -3 $O/ncu.log

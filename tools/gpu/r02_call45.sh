#!/bin/bash
# r02 call 45: per-tile clocks of the product pass (instrumented variant): where does the spread between the CTAs at the end of a pass come from?
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c45; mkdir -p $O
SKERES_LIB=$PWD/gpurun_variants/libskeres_clk.so timeout 200 python tools/tile_clock_probe.py $O/tile_clocks.npz > $O/probe.log 2>&1; tail -3 $O/probe.log

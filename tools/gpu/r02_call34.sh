#!/bin/bash
# r02 call 34: L2 residency of part of the stored Jacobian across the products of a linear solve (evict_last / evict_first copy policies)
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c34; mkdir -p $O
SKERES_L2_KEEP_MB=64 timeout 600 python -m pytest tests -m gpu -q -x -k "fused or bitwise or iterative_schur_matches" > $O/gpu_tests_keep64.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests_keep64.log
tail -n 4 $O/gpu_tests_keep64.log
fam() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['roofline']
print(f, 'value %.4g ms/step %.3f' % (d['value'], d['ms_per_step']), 'product %.4f vector %.4f frac %.4f path_frac %.4f clocks %s' % (r['product_phase_ms'], r['vector_phase_ms_per_product'], r['frac'], r['path_frac'], d['clocks']['sm_mhz']))
PY
}
for mb in -1 0 48 80 110 -1 80; do
  SKERES_L2_KEEP_MB=$mb timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/bench_keep$mb.json 2> $O/bench_keep$mb.err; fam $O/bench_keep$mb.json
done
timeout 200 python tools/batch_bench.py > $O/batch_shipped.json 2> $O/batch_shipped.err; cat $O/batch_shipped.json

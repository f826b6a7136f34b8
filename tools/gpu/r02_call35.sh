#!/bin/bash
# r02 call 35: copy policies of the product (evict_first stream / a few resident tiles) against the plain copies, same session
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c35; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "fused or bitwise or iterative_schur_matches" > $O/gpu_tests_default.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests_default.log
tail -n 4 $O/gpu_tests_default.log
fam() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['roofline']
print(f, 'value %.4g ms/step %.3f' % (d['value'], d['ms_per_step']), 'product %.4f vector %.4f frac %.4f path_frac %.4f clocks %s' % (r['product_phase_ms'], r['vector_phase_ms_per_product'], r['frac'], r['path_frac'], d['clocks']['sm_mhz']))
PY
}
i=0
for mb in -1 0 24 48 -1 0 24 48 -1 0; do
  i=$((i+1))
  SKERES_L2_KEEP_MB=$mb timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/bench_${i}_keep$mb.json 2> $O/bench_${i}_keep$mb.err; fam $O/bench_${i}_keep$mb.json
done

#!/bin/bash
# r02 call 41: per-point sums of the product as one chain per lane (three lanes per point, shuffle exchange) against one thread per point
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c41; mkdir -p $O
SKERES_LIB=$PWD/gpurun_variants/libskeres_pt3.so timeout 600 python -m pytest tests -m gpu -q -x -k "fused or bitwise or iterative_schur_matches or l2_copy or baseline_sized" > $O/gpu_tests_pt3.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests_pt3.log
tail -n 6 $O/gpu_tests_pt3.log
fam() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['roofline']
print(f, 'value %.4g ms/step %.3f' % (d['value'], d['ms_per_step']), 'product %.4f vector %.4f frac %.4f path_frac %.4f clocks %s' % (r['product_phase_ms'], r['vector_phase_ms_per_product'], r['frac'], r['path_frac'], d['clocks']['sm_mhz']))
PY
}
for i in 1 2 3; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/bench_${i}_default.json 2> $O/bench_${i}_default.err; fam $O/bench_${i}_default.json
  SKERES_LIB=$PWD/gpurun_variants/libskeres_pt3.so timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/bench_${i}_pt3.json 2> $O/bench_${i}_pt3.err; fam $O/bench_${i}_pt3.json
done

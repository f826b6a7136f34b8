#!/bin/bash
# r02 call 39: L2 policy guard test; grid barrier arriving with red.release.gpu instead of fence + atomic
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c39; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "l2_copy or fused" > $O/gpu_tests_l2.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests_l2.log
tail -n 4 $O/gpu_tests_l2.log
SKERES_LIB=$PWD/gpurun_variants/libskeres_gsred.so timeout 600 python -m pytest tests -m gpu -q -x -k "fused or bitwise" > $O/gpu_tests_gsred.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests_gsred.log
tail -n 4 $O/gpu_tests_gsred.log
fam() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['roofline']
print(f, 'value %.4g ms/step %.3f' % (d['value'], d['ms_per_step']), 'product %.4f vector %.4f frac %.4f path_frac %.4f clocks %s' % (r['product_phase_ms'], r['vector_phase_ms_per_product'], r['frac'], r['path_frac'], d['clocks']['sm_mhz']))
PY
}
for i in 1 2 3; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/bench_${i}_default.json 2> $O/bench_${i}_default.err; fam $O/bench_${i}_default.json
  SKERES_LIB=$PWD/gpurun_variants/libskeres_gsred.so timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/bench_${i}_gsred.json 2> $O/bench_${i}_gsred.err; fam $O/bench_${i}_gsred.json
done

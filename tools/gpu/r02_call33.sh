#!/bin/bash
# r02 call 33: batched curve fits with the next batch of rows loaded ahead; register caps of the set-up kernels (Jacobian evaluation / Schur set-up at 3 CTAs per SM)
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c33; mkdir -p $O
timeout 200 python tools/batch_bench.py > $O/batch_shipped.json 2> $O/batch_shipped.err; cat $O/batch_shipped.json
for v in batch_pipe4_mb5 batch_pipe4_mb4 batch_pipe2_mb5 batch_pipe4_mb3; do
  SKERES_LIB=$PWD/gpurun_variants/libskeres_$v.so timeout 200 python tools/batch_bench.py > $O/$v.json 2> $O/$v.err; cat $O/$v.json
done
fam() { python -c "
import json,sys
d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); r=d['roofline']; k=r['kernel_family_ms']; n=r['kernel_family_launches']
print('$1', 'value %.4g ms/step %.3f' % (d['value'], d['ms_per_step']), {f: round(k[f]/d['steps'],3) for f in ('evaluate_jacobian','schur_setup','evaluate_cost','back_substitute')}, 'path_frac %.4f' % r['path_frac'])"; }
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err; fam $O/bench_default.json
for v in eval3 setup3 eval1 eval2; do
  SKERES_LIB=$PWD/gpurun_variants/libskeres_$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err; fam $O/bench_$v.json
done

#!/bin/bash
# r02 call 38: where do the L2 copy policies stop paying?  Shapes with the camera count of Final-13682 and 1/8, 1/4, 1/2 of its points and observations on one GPU
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r02c38; mkdir -p $O
show() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['roofline']
    print(f, 'value %.4g ms/step %.3f' % (d['value'], d['ms_per_step']), {k:r.get(k) for k in ('vector_phase_ms_per_product','pcg_iteration_ms','frac')})
except Exception as e: print(f, 'ERR', e)
PY
}
for shape in 13682,557014,3623455 13682,1114029,7246911 13682,2228058,14493822 7000,500000,3000000; do
  for mb in 24 -1; do
    SKERES_L2_KEEP_MB=$mb timeout 600 python tools/final_scaling.py --steps 6 --warmup 2 --shape $shape > $O/final_${shape}_keep$mb.json 2> $O/final_${shape}_keep$mb.err; show $O/final_${shape}_keep$mb.json
  done
done

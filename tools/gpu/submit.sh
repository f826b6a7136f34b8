#!/bin/bash
# Development: submit a script to a B200 box, retrying while the pod has no free slot (exit code 3, nothing charged).
# usage: tools/gpu/submit.sh <log> [--gpus N] <script> [timeout_s]
log=$1; shift
gpus=""; if [ "$1" = "--gpus" ]; then gpus="--gpus $2"; shift 2; fi
script=$1; to=${2:-1800}
for i in $(seq 1 40); do
  gpurun $gpus --timeout $to -- bash $script > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "[submit] done rc=$rc after $i attempt(s)" >> $log; exit $rc; fi
  sleep 90
done
echo "[submit] gave up" >> $log

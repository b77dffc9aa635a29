#!/bin/bash
# usage: gpu_ab2.sh TAG "name bench-args..." ...   (one GPU; prints lag-0 and lag-1 us/step)
TAG=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  set -- $spec; name=$1; shift
  timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-extras --no-pipeline-baseline "$@" > gpurun_out/${TAG}_$name.log 2>&1
  python - "$TAG" "$name" <<'PY'
import json,sys
t,n=sys.argv[1],sys.argv[2]
try:
    d=json.loads([l for l in open('gpurun_out/%s_%s.log'%(t,n)).read().strip().splitlines() if l.startswith('{')][-1])
    pe=d.get('pipelined_exchange') or {}
    print(n,'steps',d['steps'],'us/step',round(d['ms_per_step']*1e3,3),'lag1',pe.get('ms_per_step') and round(pe['ms_per_step']*1e3,3))
except Exception as e:
    print(n,'failed',e); print(open('gpurun_out/%s_%s.log'%(t,n)).read()[-1500:])
PY
done

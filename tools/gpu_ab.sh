#!/bin/bash
# A/B runs of bench.py under environment switches (one GPU).  usage: gpu_ab.sh TAG "name ENV=1 ENV2=3" "name2 ..." ...
TAG=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  set -- $spec; name=$1; shift
  env "$@" timeout 300 python bench.py --steps ${STEPS:-400} --warmup 40 --no-e2e --no-cpu-baseline --no-extras --no-pipeline-baseline $BENCH_ARGS > gpurun_out/${TAG}_$name.log 2>&1
  python - "$TAG" "$name" <<'PY'
import json,sys
t,n=sys.argv[1],sys.argv[2]
try:
    d=json.loads([l for l in open('gpurun_out/%s_%s.log'%(t,n)).read().strip().splitlines() if l.startswith('{')][-1])
    print(n,'us/step',round(d['ms_per_step']*1e3,3),'frac',round(d['roofline']['frac'],4))
except Exception as e:
    print(n,'failed',e); print(open('gpurun_out/%s_%s.log'%(t,n)).read()[-1500:])
PY
done

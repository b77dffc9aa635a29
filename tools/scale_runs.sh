#!/bin/bash
# usage (GPU box with N GPUs): bash tools/scale_runs.sh N TAG
# Driver protocol (--steps 20 --warmup 5) at N GPUs: (a) weak scaling, B = 64 per GPU (the headline of SCALE_rNN.json),
# (b) BASELINE configs[4]: the strong split of a 512-sequence batch, B = 512 / N per GPU, (c) 400 steps for reference.
N=$1; TAG=${2:-r2}
mkdir -p gpurun_out
run() { name=$1; shift
  if [ "$N" = "1" ]; then timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/${TAG}_${name}_n$N.log 2>&1
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/${TAG}_${name}_n$N.log 2>&1; fi
  python - "$name" "$N" "$TAG" <<'PY'
import json,sys
n,N,tag=sys.argv[1:4]
try:
    d=json.loads([l for l in open('gpurun_out/%s_%s_n%s.log'%(tag,n,N)).read().strip().splitlines() if l.startswith('{')][-1])
    pe=d.get('pipelined_exchange') or {}
    print(n, 'N',N,'B/gpu',d['config']['batch_per_gpu'],'us/step', round(d['ms_per_step']*1e3,2), 'value', round(d['value']/1e6,2),'M tok/s', 'frac',round(d['roofline']['frac'],3),'lag1', pe.get('ms_per_step') and round(pe['ms_per_step']*1e3,2), 'e2e ms', d.get('e2e',{}).get('ms_per_step'))
except Exception as e:
    print(n,'failed',e); print(open('gpurun_out/%s_%s_n%s.log'%(tag,n,N)).read()[-2500:])
PY
}
run weak --steps 20 --warmup 5 --no-cpu-baseline --no-extras
run weak400 --steps 400 --warmup 40 --no-cpu-baseline --no-extras --no-e2e
if [ "$N" != "1" ]; then
  run strong512 --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-e2e --batch $((512 / N))
fi

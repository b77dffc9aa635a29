#!/bin/bash
# usage (box with N GPUs): bash tools/gpu_n.sh N TAG  -- exchange correctness check + driver-protocol weak run (+ 400 steps)
N=$1; TAG=${2:-r2u}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/check_peer_exchange.py > gpurun_out/${TAG}_peer_n$N.log 2>&1; echo "check_peer_exchange rc=$?"; tail -2 gpurun_out/${TAG}_peer_n$N.log | cut -c1-600
run() { name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/${TAG}_${name}_n$N.log 2>&1
  python - "$name" "$N" "$TAG" <<'PY'
import json,sys
n,N,tag=sys.argv[1:4]
try:
    d=json.loads([l for l in open('gpurun_out/%s_%s_n%s.log'%(tag,n,N)).read().strip().splitlines() if l.startswith('{')][-1])
    pe=d.get('pipelined_exchange') or {}
    print(n, 'N',N,'B/gpu',d['config']['batch_per_gpu'],'us/step', round(d['ms_per_step']*1e3,2), 'value', round(d['value']/1e6,2),'M tok/s', 'lag1', pe.get('ms_per_step') and round(pe['ms_per_step']*1e3,2))
except Exception as e:
    print(n,'failed',e); print(open('gpurun_out/%s_%s_n%s.log'%(tag,n,N)).read()[-2500:])
PY
}
run weak --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-e2e --no-pipeline-baseline
# run weak400 --steps 400 --warmup 40 --no-cpu-baseline --no-extras --no-e2e --no-pipeline-baseline
if [ -n "$WITH_400" ]; then run weak400 --steps 400 --warmup 40 --no-cpu-baseline --no-extras --no-e2e --no-pipeline-baseline; fi

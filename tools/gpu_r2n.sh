#!/bin/bash
# usage: gpu_r2n.sh N
N=$1
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/r2n_${name}_n$N.log 2>&1
  python - "$name" "$N" <<'PY'
import json,sys
n,N=sys.argv[1],sys.argv[2]
try:
    d=json.loads([l for l in open('gpurun_out/r2n_%s_n%s.log'%(n,N)).read().strip().splitlines() if l.startswith('{')][-1])
    print(n, 'N',N,'us/step', round(d['ms_per_step']*1e3,2), 'value', round(d['value']/1e6,2),'M tok/s', 'lag1', d.get('pipelined_exchange',{}) and round(d['pipelined_exchange']['ms_per_step']*1e3,2), 'e2e', d.get('e2e',{}).get('ms_per_step'))
except Exception as e:
    print(n,'failed',e); print(open('gpurun_out/r2n_%s_n%s.log'%(n,N)).read()[-2500:])
PY
}
run driver --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-e2e
run driver_b --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-e2e
run s400 --steps 400 --warmup 40 --no-cpu-baseline --no-extras --no-e2e
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2n_pytest.log 2>&1; tail -3 gpurun_out/r2n_pytest.log

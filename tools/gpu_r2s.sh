#!/bin/bash
# A/B of the tail switches of the fused step (one GPU): SD_B200_STEP_TAIL_ORDER / _TAIL_PEEK / _LAG_PCT
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 400 --warmup 40 --no-e2e --no-cpu-baseline --no-extras --no-pipeline-baseline > gpurun_out/r2s_$name.log 2>&1
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open('gpurun_out/r2s_%s.log'%n).read().strip().splitlines() if l.startswith('{')][-1])
    print(n,'us/step',round(d['ms_per_step']*1e3,3),'frac',round(d['roofline']['frac'],4))
except Exception as e:
    print(n,'failed',e); print(open('gpurun_out/r2s_%s.log'%n).read()[-1500:])
PY
}
run base A=0
run order1 SD_B200_STEP_TAIL_ORDER=1
run peek SD_B200_STEP_TAIL_PEEK=1
run order1_peek SD_B200_STEP_TAIL_ORDER=1 SD_B200_STEP_TAIL_PEEK=1
run order1_peek_lag160 SD_B200_STEP_TAIL_ORDER=1 SD_B200_STEP_TAIL_PEEK=1 SD_B200_STEP_LAG_PCT=160
run order1_peek_lag200 SD_B200_STEP_TAIL_ORDER=1 SD_B200_STEP_TAIL_PEEK=1 SD_B200_STEP_LAG_PCT=200
run order1_peek_lag300 SD_B200_STEP_TAIL_ORDER=1 SD_B200_STEP_TAIL_PEEK=1 SD_B200_STEP_LAG_PCT=300
run base2 A=0
SD_B200_STEP_TAIL_ORDER=1 SD_B200_STEP_TAIL_PEEK=1 timeout 600 python -m pytest tests/test_verify_step_gpu.py tests/test_device_loop.py tests/test_canaries_gpu.py -q -m gpu > gpurun_out/r2s_pytest.log 2>&1; echo "pytest(order1,peek) rc=$?"; tail -2 gpurun_out/r2s_pytest.log
timeout 600 python -m pytest tests/test_verify_step_gpu.py tests/test_device_loop.py -q -m gpu > gpurun_out/r2s_pytest0.log 2>&1; echo "pytest(default) rc=$?"; tail -2 gpurun_out/r2s_pytest0.log
SD_B200_STEP_TAIL_ORDER=1 SD_B200_STEP_TAIL_PEEK=1 timeout 300 python tools/trace_step.py > gpurun_out/r2s_trace.log 2>&1; tail -1 gpurun_out/r2s_trace.log

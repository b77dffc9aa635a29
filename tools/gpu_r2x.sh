#!/bin/bash
# Balanced (persistent) attention: parity tests, then both forms timed by bench.py's isolation loops.
TAG=${1:-r2x}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_attention_gpu.py -x -q > $OUT/pytest_attn_$TAG.log 2>&1; rc=$?
echo "pytest attention rc=$rc $(tail -1 $OUT/pytest_attn_$TAG.log)"
if [ $rc -ne 0 ]; then grep -E "^E|Error|FAILED" $OUT/pytest_attn_$TAG.log | head -30; fi
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-pipeline-baseline > $OUT/bench_$TAG.log 2>&1; echo "bench rc=$?"
python - "$OUT/bench_$TAG.log" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print('step us', d['ms_per_step']*1e3)
    for k in ('roofline_attn','roofline_attn_d64'):
        print(k, d[k].get('form'), d[k].get('us_per_launch_by_form'), round(d[k]['frac'],3))
except Exception as e:
    print('parse failed', e); print(open(sys.argv[1]).read()[-2000:])
PY

#!/bin/bash
# Diagnostic (GPU box): is the fused step limited by the page spread of its KV stores?  Builds a throw-away copy of the
# library whose append folds every destination into 1 MB (results are WRONG; timing only) and times both.
set -u
mkdir -p gpurun_out /tmp/diag
LIB=llm-inference-lab_b200/lib/libsd_verify_b200.so
cp $LIB /tmp/diag/good.so
run() { timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-extras --steps 400 --warmup 40 "$@" 2>&1 | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2),'us/step  kv alone',round(d['roofline_kv']['us_per_launch'],2))"; }
echo "normal:"; run
for f in sd_api sd_attn sd_exchange sd_kv sd_sample sd_verify; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -DSD_DIAG_ALIAS_DST -I include -c llm-inference-lab_b200/csrc/$f.cu -o /tmp/diag/$f.o &
done; wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o /tmp/diag/alias.so /tmp/diag/*.o
cp /tmp/diag/alias.so $LIB
echo "aliased destinations (1 MB):"; run
cp /tmp/diag/good.so $LIB

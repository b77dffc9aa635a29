#!/bin/bash
# ncu --set full of the D=64 (Llama-3.2-1B geometry) attention kernel, standard and balanced form
TAG=${1:-r2y}
OUT=gpurun_out
mkdir -p $OUT
SHORT="python bench.py --steps 8 --warmup 4 --group 4 --no-e2e --no-cpu-baseline --no-pipeline-baseline --no-gate"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:sd::verify_attn_kernel<__nv_bfloat16, .int.64" -s 10 -c 2 -f -o $OUT/prof_attn64_std_$TAG $SHORT > $OUT/ncu_attn64_std_$TAG.log 2>&1; echo "ncu std rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:sd::verify_attn_balanced_kernel<__nv_bfloat16, .int.64" -s 10 -c 2 -f -o $OUT/prof_attn64_bal_$TAG $SHORT > $OUT/ncu_attn64_bal_$TAG.log 2>&1; echo "ncu bal rc=$?"
ls -la $OUT/*$TAG* | awk '{print $5, $9}'

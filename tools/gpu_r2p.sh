#!/bin/bash
mkdir -p gpurun_out
bash tools/capture_profiles.sh r2a 2>&1 | tail -15
timeout 900 python scripts/microbench_verify.py --markdown gpurun_out/r2_microbench.md > gpurun_out/r2_microbench.log 2>&1
echo "microbench rc=$?"; tail -3 gpurun_out/r2_microbench.log | cut -c1-300

#!/bin/bash
TAG=${1:-r2b}
mkdir -p gpurun_out
bash tools/capture_profiles.sh $TAG 2>&1 | tail -14
timeout 600 python bench.py --steps 2000 --warmup 40 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/bench2000_$TAG.log 2>&1; grep '^{' gpurun_out/bench2000_$TAG.log | cut -c1-260
timeout 600 python bench.py --steps 400 --warmup 40 --no-e2e --no-cpu-baseline --no-extras --fused-mode 1 > gpurun_out/bench_mode1_$TAG.log 2>&1; grep '^{' gpurun_out/bench_mode1_$TAG.log | cut -c1-260
timeout 600 python bench.py --steps 400 --warmup 40 --no-e2e --no-cpu-baseline --no-extras --no-fuse > gpurun_out/bench_nofuse_$TAG.log 2>&1; grep '^{' gpurun_out/bench_nofuse_$TAG.log | cut -c1-260
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference_$TAG.log 2>&1; tail -1 gpurun_out/bench_reference_$TAG.log | cut -c1-260
timeout 300 python tools/trace_step.py > gpurun_out/trace_step_$TAG.log 2>&1; tail -1 gpurun_out/trace_step_$TAG.log | cut -c1-400
timeout 300 python tools/host_overhead.py > gpurun_out/host_overhead_$TAG.log 2>&1; cat gpurun_out/host_overhead_$TAG.log | tr -d '\n' | cut -c1-900; echo
timeout 900 python scripts/microbench_verify.py --markdown gpurun_out/microbench_$TAG.md > gpurun_out/microbench_$TAG.log 2>&1; echo "microbench rc=$?"
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-300

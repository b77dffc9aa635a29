#!/bin/bash
# Final one-GPU pass of the round: parity suite, driver-protocol bench line (all legs), reference arm, step timeline
# with and without the (loopback) exchange, smoke.
TAG=${1:-r2f}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/pytest_gpu_$TAG.log)"
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.log 2>&1; echo "bench rc=$?"; grep '^{' $OUT/bench_$TAG.log | tail -1 | cut -c1-260
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference_$TAG.log 2>&1; tail -1 $OUT/bench_reference_$TAG.log | cut -c1-200
timeout 120 python tools/trace_step.py --reps 4 > $OUT/trace_step_$TAG.jsonl 2>$OUT/trace_step_$TAG.err; echo "trace rc=$?"; tail -1 $OUT/trace_step_$TAG.jsonl | cut -c1-300
timeout 120 python tools/trace_step.py --reps 4 --loopback > $OUT/trace_step_loopback_$TAG.jsonl 2>$OUT/trace_step_loopback_$TAG.err; echo "trace loopback rc=$?"; tail -1 $OUT/trace_step_loopback_$TAG.jsonl | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | cut -c1-300

#!/bin/bash
# N-GPU per-CTA timeline of the step with the peer-memory exchange (lock-step, lag 0)
N=${1:-2}; TAG=${2:-r2g}
mkdir -p gpurun_out
timeout 50 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/trace_step.py --exchange --xlag 0 --reps 4 > gpurun_out/trace_step_n${N}_lag0_$TAG.jsonl 2> gpurun_out/trace_step_n${N}_$TAG.err; echo "rc=$?"
grep '^{' gpurun_out/trace_step_n${N}_lag0_$TAG.jsonl | tail -2 | cut -c1-330

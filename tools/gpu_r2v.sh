#!/bin/bash
# One-GPU evidence pass for the tree at HEAD (round 2, last sessions): parity suite, driver-protocol bench line,
# logits-prefetch A/B of the step kernel, launch list + ncu --set full of the dominant kernel, smoke.
TAG=${1:-r2v}
OUT=gpurun_out
mkdir -p $OUT
SHORT="python bench.py --steps 8 --warmup 4 --group 4 --no-e2e --no-cpu-baseline --no-pipeline-baseline --no-gate"
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/pytest_gpu_$TAG.log)"
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.log 2>&1; echo "bench rc=$?"; grep '^{' $OUT/bench_$TAG.log | tail -1 | cut -c1-300
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference_$TAG.log 2>&1; tail -1 $OUT/bench_reference_$TAG.log | cut -c1-200
bash tools/gpu_ab2.sh $TAG "pf0 --steps 400 --warmup 40" "pf0b --steps 2000 --warmup 40"
for pf in 4 16 72 148; do SD_B200_STEP_PF_AHEAD=$pf bash tools/gpu_ab2.sh $TAG "pf$pf --steps 400 --warmup 40"; done
timeout 200 $SHORT > $OUT/ncu_plain_$TAG.log 2>&1 && {
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv $SHORT > $OUT/ncu_list_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:verify_step_kernel" -s 2 -c 2 -f -o $OUT/prof_verify_step_$TAG $SHORT > $OUT/ncu_full_verify_step_$TAG.log 2>&1; echo "ncu full rc=$?"
}
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | cut -c1-300

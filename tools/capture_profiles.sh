#!/usr/bin/env bash
# One-GPU evidence pass, to be run on the GPU box (under gpurun) from the repository root:
#
#     gpurun --timeout 1500 -- 'bash tools/capture_profiles.sh r2a'
#
# 1. the GPU parity suite, 2. one plain bench line under the driver's protocol (the only numbers that may be quoted),
# 3. the ncu launch list of a short bench (shares only: per-launch times under ncu are cold-cache and serialised),
# 4. one `ncu --set full` capture per dominant kernel (2 launches each, after skipping the warm-up launches).
# Everything lands in gpurun_out/; back in the build container summarise with
#     python tools/ncu_summary.py --round <tag> --launches gpurun_out/launches_<tag>.csv gpurun_out/prof_*_<tag>.ncu-rep
# Each ncu step runs only if the plain command before it exited 0.  The profiled command runs WITHOUT the start gate:
# under ncu every launch is synchronous, so a gate waiting for "everything behind me is enqueued" could never open.
set -u
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p "$OUT"
SHORT="python bench.py --steps 8 --warmup 4 --group 4 --no-e2e --no-cpu-baseline --no-gate"

timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu_$TAG.log" 2>&1
echo "pytest rc=$? $(tail -1 "$OUT/pytest_gpu_$TAG.log")"

timeout 900 python bench.py --steps 20 --warmup 5 > "$OUT/bench_$TAG.log" 2>&1 || { echo "bench failed"; tail -5 "$OUT/bench_$TAG.log"; exit 1; }
grep '^{' "$OUT/bench_$TAG.log" | tail -1 | cut -c1-400

timeout 300 $SHORT > "$OUT/ncu_plain_$TAG.log" 2>&1 || { echo "short bench failed"; tail -5 "$OUT/ncu_plain_$TAG.log"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches_$TAG.csv" \
    $SHORT > "$OUT/ncu_list_$TAG.log" 2>&1
echo "launch list rc=$?"

for K in verify_step_kernel verify_bulk_kernel kv_append_col_kernel verify_attn_kernel; do
    short=${K%_kernel}
    timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$K" -s 2 -c 2 -f \
        -o "$OUT/prof_${short}_$TAG" $SHORT > "$OUT/ncu_full_${short}_$TAG.log" 2>&1
    echo "ncu --set full $K rc=$?"
done
ls -la "$OUT"/*_"$TAG"* 2>/dev/null | awk '{print $5, $9}'

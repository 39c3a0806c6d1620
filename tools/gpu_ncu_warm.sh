#!/bin/bash
# Warm-cache (no flush between replays) per-launch durations of one timed bench step.  Usage: bash tools/gpu_ncu_warm.sh <tag>
TAG=${1:-w}
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --ncu > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max --clock-control none --cache-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_warm_$TAG.csv python bench.py --steps 3 --warmup 3 --ncu > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/launches_warm_$TAG.csv

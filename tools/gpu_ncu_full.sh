#!/bin/bash
# ncu --set full capture of selected kernels inside one timed bench step.
# Usage: bash tools/gpu_ncu_full.sh <tag> <kernel-regex> <skip> <count>
TAG=$1; KRE=$2; SKIP=${3:-0}; CNT=${4:-3}
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --ncu > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$KRE -s $SKIP -c $CNT \
    -f -o gpurun_out/prof_$TAG python bench.py --steps 3 --warmup 3 --ncu > gpurun_out/ncufull_$TAG.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncufull_$TAG.log; ls -la gpurun_out/prof_$TAG.ncu-rep

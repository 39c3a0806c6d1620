#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, then the ncu launch list of one timed bench step.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi -L; nproc
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_$TAG.log
python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke_$TAG.log
python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_$TAG.err | tee gpurun_out/bench_$TAG.json
tail -5 gpurun_out/bench_$TAG.err
python bench.py --steps 3 --warmup 3 --ncu > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --ncu > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_$TAG.log; wc -l gpurun_out/launches_$TAG.csv

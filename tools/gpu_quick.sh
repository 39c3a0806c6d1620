#!/bin/bash
# Quick GPU pass: parity tests, bench line, in-graph attribution.  Usage: bash tools/gpu_quick.sh [tag]
TAG=${1:-q}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_$TAG.err | tee gpurun_out/bench_$TAG.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'roof', round(r['frac'],3))
print({k: round(v,3) for k,v in r['classes_ms_per_step'].items()})"
tail -3 gpurun_out/bench_$TAG.err
python tools/ablate.py 2>&1 | tee gpurun_out/ablate_$TAG.log

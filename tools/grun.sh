#!/bin/bash
# Build both libraries, then run a command on the GPU box:  tools/grun.sh [--gpus N] [--timeout S] -- '<command>'
set -e
cd "$(dirname "$0")/.."
make -s -C styletts-zs_b200/csrc libstz.so libstz_trace.so 2>&1 | grep -E "error|warning" || true
exec /usr/local/graft/bin/gpurun "$@"

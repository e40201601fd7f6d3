#!/bin/bash
# build the CUDA library in-tree, then run a command on the B200 box:  scripts/gpu.sh [timeout_s] 'command'
set -e
cd "$(dirname "$0")/.."
T=600
if [[ "$1" =~ ^[0-9]+$ ]]; then T=$1; shift; fi
python -m liana_py_b200.build > /dev/null
exec /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"

#!/bin/bash
# Rebuild libhvla.so here (a stale .so would travel to the GPU box as it is), then run a command on a B200 through gpurun.
#   tools/gpu.sh [--timeout S] [--gpus N] -- '<command>'
set -e
cd "$(dirname "$0")/.."
python __graft_entry__.py build | tail -1
exec /usr/local/graft/bin/gpurun "$@"

#!/bin/bash
# gpurun with the guard this repository needs: the GPU box runs the PREBUILT library, so everything is
# rebuilt here first (a stale .so once sent a fixed bug back to the GPU).   tools/gpurun.sh [gpurun options] -- 'command'
set -e
cd "$(dirname "$0")/.."
make -j8 > /tmp/gpurun_make.log 2>&1 || { tail -20 /tmp/gpurun_make.log; exit 1; }
make -C tests/integration > /dev/null 2>&1 || true
exec /usr/local/graft/bin/gpurun "$@"

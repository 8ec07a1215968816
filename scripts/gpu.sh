#!/bin/bash
# Build everything (CUDA library + oracle), then run a command on a B200 through gpurun.
#   scripts/gpu.sh [--gpus N] TIMEOUT_SECONDS 'command'
set -e
cd "$(dirname "$0")/.."
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
make -C adaptive_mcmc_b200/csrc -j"$(nproc)" 2>&1 | grep -E "error|Error|undefined" && exit 1
make -C oracle >/dev/null
exec /usr/local/graft/bin/gpurun $GP --timeout "$T" -- "$@"

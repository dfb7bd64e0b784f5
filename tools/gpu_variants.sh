#!/bin/bash
# Times experiment builds (tools/variants.sh) of one configuration, one process per library.
# usage: tools/gpu_variants.sh "<kbench args>" lib1 lib2 ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ARGS=$1; shift
for lib in "$@"; do
  timeout 300 python tools/kbench.py $ARGS --lib $lib 2>&1 | tail -1
done | tee -a gpurun_out/variants.log

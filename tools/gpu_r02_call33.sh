#!/bin/bash
# Round 2, GPU call 33 (1 GPU): pageable end-to-end rate against the size of the staging-copy thread pool (tools/pageable_sweep.py).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket|NUMA node\(s\)" | head -6
  timeout 400 python tools/pageable_sweep.py --threads default,4,8,12,16,8
} 2>&1 | tee gpurun_out/r02ae_call33.log

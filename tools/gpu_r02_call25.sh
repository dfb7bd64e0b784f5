#!/bin/bash
# Round 2, GPU call 25 (1 GPU): bench of the tree as it ships (both arms, wall time of each).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02x
{
  echo "== reference arm (driver's command line)"
  s=$(date +%s); timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > ${O}_ref.json 2>/dev/null; echo "rc=$? wall=$(( $(date +%s) - s )) s"; cut -c1-600 ${O}_ref.json
  echo "== bench (N=1)"
  s=$(date +%s); timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$? wall=$(( $(date +%s) - s )) s"; cut -c1-300 ${O}_bench.json; tail -2 ${O}_bench.err
} 2>&1 | tee ${O}_call25.log

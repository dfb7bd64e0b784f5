#!/bin/bash
# Round 2, GPU call 26 (1 GPU): the tree rebuilt from a clean checkout (the container was re-created: every .so is a fresh build of the
# committed sources) -- whole GPU suite, smoke, bench of both arms with the driver's command lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02y
{
  nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
  echo "== GPU suite"
  s=$(date +%s); timeout 1500 python -m pytest tests -m gpu -q -x > ${O}_pytest_gpu.log 2>&1; echo "rc=$? wall=$(( $(date +%s) - s )) s"; tail -4 ${O}_pytest_gpu.log
  echo "== smoke"
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
  echo "== reference arm (driver's command line)"
  s=$(date +%s); timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > ${O}_ref.json 2>/dev/null; echo "rc=$? wall=$(( $(date +%s) - s )) s"; cut -c1-600 ${O}_ref.json
  echo "== bench (N=1)"
  s=$(date +%s); timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$? wall=$(( $(date +%s) - s )) s"; cut -c1-400 ${O}_bench.json; tail -2 ${O}_bench.err
} 2>&1 | tee ${O}_call26.log

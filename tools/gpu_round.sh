#!/bin/bash
# One GPU-box visit of the development loop: variant timings, then the GPU suite.  Results land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
{
  echo "== kbench C2 variants"
  timeout 300 python tools/kbench.py --lib build/variants/static/libnflgpu.so --lib build/variants/dyn1/libnflgpu.so --lib build/variants/dyn2/libnflgpu.so --lib build/variants/static/libnflgpu.so --lib build/variants/dyn1/libnflgpu.so --lib build/variants/dyn2/libnflgpu.so
  for v in static_full dyn2_full; do
    echo "== kbench_all $v"
    NFLGPU_LIB=build/variants/$v/libnflgpu.so timeout 600 python tools/kbench_all.py
  done
  echo "== kbench_all default (dyn1)"
  timeout 600 python tools/kbench_all.py
} > gpurun_out/variants.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
cat gpurun_out/variants.log

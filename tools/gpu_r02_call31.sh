#!/bin/bash
# Round 2, GPU call 31 (1 GPU): CTA-local dynamic unit walk for N = 1024 x 64-bit (-DNFLGPU_LOCAL=1: the unit slots of a CTA draw the CTA's
# static share of units from a shared-memory counter) = local10, the global dynamic walk on the current geometry (-DNFLGPU_DYNAMIC=1) = dyn10,
# against the tree's static walk = base10.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02ad
{
  echo "== whole-batch parity of the local walk (batches 1, 37, 4096, 4099; three launches each)"
  timeout 300 python tools/check_variant.py --lib build/variants/local10/libnflgpu.so 2>&1 | tail -5
  echo "== C2 u64 N=1024 M=4 batch=4096"
  for v in base10 local10 dyn10 base10 local10 dyn10; do timeout 300 python tools/kbench.py --bits 64 --degree 1024 --nmoduli 4 --batch 4096 --lib build/variants/$v/libnflgpu.so 2>&1 | tail -1; done
  echo "== u64 N=1024 M=4 batch=1000 (ragged: 4000 units over 2072 slots)"
  for v in base10 local10 base10 local10; do timeout 300 python tools/kbench.py --bits 64 --degree 1024 --nmoduli 4 --batch 1000 --lib build/variants/$v/libnflgpu.so 2>&1 | tail -1; done
} 2>&1 | tee ${O}_call31.log

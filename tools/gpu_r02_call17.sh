#!/bin/bash
# Round 2, GPU call 17 (1 GPU): TMA bulk staging of the coefficient slab (-DNFLGPU_TMA_SLAB experiment) against the tree, N = 1024 x 64-bit.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02q
{
  echo "== -DNFLGPU_TMA_SLAB (tma10) vs tree (base10), C2 u64 N=1024 M=4 batch=4096 (forward only differs)"
  for v in base10 tma10 base10 tma10 base10 tma10; do timeout 300 python tools/kbench.py --bits 64 --degree 1024 --nmoduli 4 --batch 4096 --lib build/variants/$v/libnflgpu.so 2>&1 | tail -1; done
  echo "== ragged batch (37 polys) and one poly through the experiment build: parity of the bulk-copy path"
  for b in 37 1; do timeout 300 python tools/kbench.py --bits 64 --degree 1024 --nmoduli 4 --batch $b --iters 3 --lib build/variants/tma10/libnflgpu.so 2>&1 | tail -1; done
} 2>&1 | tee ${O}_call17.log

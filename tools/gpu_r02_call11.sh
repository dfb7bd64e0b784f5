#!/bin/bash
# Round 2, GPU call 11 (1 GPU): where the per-chunk cost of the host pipeline comes from (pure copies, no-kernel build, larger chunks);
# XOR-swizzled tile of the 32-bit N = 4096 kernels against the padded one.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02k
{
  echo "== chunked copies without the library (tools/copy_chunks.py)"
  timeout 300 python tools/copy_chunks.py 2>&1 | tee ${O}_copy_chunks.txt
  echo "== e2e sweep, larger chunks"
  timeout 600 python tools/e2e_sweep.py 16:3 16:4 32:3 32:4 64:3 2>&1 | tee ${O}_e2e_sweep.txt
  echo "== e2e sweep, build without kernels"
  E2E_LIB=build/variants/nokernel/libnflgpu.so timeout 600 python tools/e2e_sweep.py 4:8 16:4 2>&1 | tee -a ${O}_e2e_sweep.txt
  echo "== swizzled tile, C4 shape (u32 N=4096 M=14), batch 2048 and 8192"
  for b in 2048 8192; do
    for v in noswz12 swz12 noswz12 swz12; do
      timeout 300 python tools/kbench.py --bits 32 --degree 4096 --nmoduli 14 --batch $b --lib build/variants/$v/libnflgpu.so 2>&1 | tail -1
    done
  done | tee ${O}_swizzle.txt
  echo "== parity of the swizzled kernels"
  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sizes or full or C4 or fixtures" 2>&1 | tail -4
} 2>&1 | tee ${O}_call11.log

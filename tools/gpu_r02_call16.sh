#!/bin/bash
# Round 2, GPU call 16 (1 GPU): adjacent-column pass 0 (128-bit HBM loads / stores, NttCfg::ADJ) against the strided 8-byte form (-DNFLGPU_ADJ=0).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02p
{
  echo "== NttCfg::ADJ (adj) vs -DNFLGPU_ADJ=0 (noadj)"
  kb() { for v in noadj$1 adj$1 noadj$1 adj$1; do timeout 300 python tools/kbench.py $2 --lib build/variants/$v/libnflgpu.so 2>&1 | tail -1; done; }
  echo "# C2 u64 N=1024 M=4 batch=4096";   kb 10 "--bits 64 --degree 1024 --nmoduli 4 --batch 4096"
  echo "# u64 N=2048 M=4 batch=2048";      kb 11 "--bits 64 --degree 2048 --nmoduli 4 --batch 2048"
  echo "# C3 u64 N=16384 M=8 batch=512";   kb 14 "--bits 64 --degree 16384 --nmoduli 8 --batch 512"
  echo "== parity (tree = adj)"
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_round2.py -m gpu -x -q -k "sizes or full or fixtures or appendix or baseline or edge or host" 2>&1 | tail -3
} 2>&1 | tee ${O}_call16.log

#!/bin/bash
# Round 2, GPU call 4: whole GPU suite on the current tree, bench line, ablation / geometry variants of the headline kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02d
{
  echo "== variants N=1024 u64 M=4 batch 4096 (v_base = tree: 1024 threads x 1 CTA)"
  for v in v_base v_abl1 v_abl2 v_abl4 v_abl8 v_abl15 v_dyn v_nopf v_plain v_t768 v_t896 v_base; do
    timeout 300 python tools/kbench.py --bits 64 --degree 1024 --nmoduli 4 --batch 4096 --lib build/variants/$v/libnflgpu.so 2>&1 | tail -1
  done
  echo "== bench (N=1)"
  timeout 900 python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$?"; cut -c1-3000 ${O}_bench.json; tail -5 ${O}_bench.err
  echo "== GPU suite"
  timeout 1800 python -m pytest tests -m gpu -x -q > ${O}_pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 ${O}_pytest_gpu.log
} 2>&1 | tee ${O}_call4.log

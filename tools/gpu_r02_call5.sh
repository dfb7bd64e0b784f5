#!/bin/bash
# Round 2, GPU call 5: ablations / geometry of the headline kernel, cluster kernels vs the round-1 split path, whole GPU suite.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02e
{
  echo "== variants N=1024 u64 M=4 batch 4096 (v_base = tree: 1024 threads x 1 CTA; abl = timing-only ablations, results wrong by design)"
  for v in v_base v_abl1 v_abl2 v_abl4 v_abl8 v_abl15 v_dyn v_nopf v_plain v_t768 v_t896 v_base; do
    timeout 300 python tools/kbench.py --bits 64 --degree 1024 --nmoduli 4 --batch 4096 --lib build/variants/$v/libnflgpu.so 2>&1 | tail -1
  done
  echo "== cluster kernels (tree) vs global-memory pass + tile kernel (NFLGPU_NO_CLUSTER=1)"
  for cfg in "64 32768 2 256" "64 32768 4 512" "64 65536 2 128"; do
    set -- $cfg
    echo "# u$1 N=$2 M=$3 batch=$4"
    timeout 300 python tools/kbench.py --bits $1 --degree $2 --nmoduli $3 --batch $4 2>&1 | tail -1
    NFLGPU_NO_CLUSTER=1 timeout 300 python tools/kbench.py --bits $1 --degree $2 --nmoduli $3 --batch $4 2>&1 | tail -1
  done
  echo "== GPU suite"
  timeout 2400 python -m pytest tests -m gpu -x -q > ${O}_pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 ${O}_pytest_gpu.log
  echo "== bench (N=1)"
  timeout 900 python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$?"; cut -c1-1500 ${O}_bench.json; tail -5 ${O}_bench.err
} 2>&1 | tee ${O}_call5.log

#!/bin/bash
# Round 2, GPU call 14 (1 GPU): software pipelining across units (-DNFLGPU_PIPE=1: inverse cp.async copy-in under the last pass, forward pass-0
# loads under the copy-out) against the tree, per size; e2e with 32 MiB asynchronous chunks and the copy-thread pool.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02n
{
  echo "== NFLGPU_PIPE=1 vs tree"
  kb() { for v in base$1 pipe$1 base$1 pipe$1; do timeout 300 python tools/kbench.py $2 --lib build/variants/$v/libnflgpu.so 2>&1 | tail -1; done; }
  echo "# C2 u64 N=1024 M=4 batch=4096";   kb 10 "--bits 64 --degree 1024 --nmoduli 4 --batch 4096"
  echo "# C4 u32 N=4096 M=14 batch=2048";  kb 12 "--bits 32 --degree 4096 --nmoduli 14 --batch 2048"
  echo "# u64 N=4096 M=4 batch=1024";      kb 12 "--bits 64 --degree 4096 --nmoduli 4 --batch 1024"
  echo "# C5 u64 N=8192 M=6 batch=2048";   kb 13 "--bits 64 --degree 8192 --nmoduli 6 --batch 2048"
  echo "# C3 u64 N=16384 M=8 batch=1024";  kb 14 "--bits 64 --degree 16384 --nmoduli 8 --batch 1024"
  echo "== bench (headline only): e2e with 32 MiB async chunks, pageable path with the copy pool"
  timeout 900 python bench.py --headline-only --no-cpu > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$?"
  python -c "
import json;d=json.load(open('${O}_bench.json'));e=d['e2e']
print('e2e',e['value'],'one_wait',e['one_wait_per_step']['value'],'blocking',e['blocking_calls']['value'],'pageable',e['pageable']['value'],'registered',e['pageable_registered']['value'],'ceiling',e['copy_only_ceiling']['value'],'lat',e['single_poly_latency_us']['median'])"
  tail -3 ${O}_bench.err
} 2>&1 | tee ${O}_call14.log

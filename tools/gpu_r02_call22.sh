#!/bin/bash
# Round 2, GPU call 22 (1 GPU): cluster inverse kernel with the next unit's cp.async copy-in under pass 0 (tree) against -DNFLGPU_PIPE=0, N = 2^15 x 64-bit;
# cluster tests + racecheck / memcheck of the launch stress (N = 32768 shape with several units per cluster).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02u
{
  echo "== cluster inverse pipelining: nopipe15 vs pipe15 (forward kernels identical)"
  for args in "--nmoduli 2 --batch 256" "--nmoduli 4 --batch 512"; do
    echo "# u64 N=32768 $args"
    for v in nopipe15 pipe15 nopipe15 pipe15; do timeout 300 python tools/kbench.py --bits 64 --degree 32768 $args --lib build/variants/$v/libnflgpu.so 2>&1 | tail -1; done
  done
  echo "== cluster tests (tree)"
  timeout 900 python -m pytest tests -m gpu -x -q -k "cluster or live_reference or sizes" 2>&1 | tail -3
  echo "== reference programs at 32768 (tree)"
  timeout 900 python -m pytest tests/test_reference_programs.py -m gpu -x -q 2>&1 | tail -2
  echo "== racecheck / memcheck, N = 32768 with more units than clusters"
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 tests/cpp/sched_stress 8 32768 2>&1 | tail -4
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 tests/cpp/sched_stress 8 32768 2>&1 | tail -4
} 2>&1 | tee ${O}_call22.log

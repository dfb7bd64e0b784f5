#!/bin/bash
# Round 2, GPU call 23 (1 GPU): small host calls through mapped pinned memory (no copy engines): single-poly latency with / without, host-buffer tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02v
{
  echo "== single-poly latency (tools/e2e_sweep.py prints it last): direct path off (NFLGPU_HOST_SMALL_KIB=0) / on (default 128)"
  for k in 0 128 0 128; do echo "-- NFLGPU_HOST_SMALL_KIB=$k"; NFLGPU_HOST_SMALL_KIB=$k timeout 300 python tools/e2e_sweep.py 16:4 2>&1 | tail -1; done
  echo "== latency by shape (python tools/host_latency.py)"
  for k in 0 128; do echo "-- NFLGPU_HOST_SMALL_KIB=$k"; NFLGPU_HOST_SMALL_KIB=$k timeout 300 python tools/host_latency.py 2>&1 | tail -8; done
  echo "== host-buffer tests, drop-in programs, reference programs"
  timeout 1200 python -m pytest tests -m gpu -x -q -k "host or dropin or reference_programs or round2 or smoke" 2>&1 | tail -3
  echo "== memcheck over the drop-in test program (single-poly calls) "
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_cpp_dropin.py -m gpu -x -q 2>&1 | tail -4
} 2>&1 | tee ${O}_call23.log

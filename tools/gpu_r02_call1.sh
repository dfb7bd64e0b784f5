#!/bin/bash
# Round 2, GPU call 1: the checks round 1 left CPU-only (reference demo programs, sanitizers on the final kernels) + pipe micro-benchmarks.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02a
{
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
  echo "== pipe micro-benchmarks"
  timeout 300 tools/ubench/pipemix
  echo "== reference demo programs against the drop-in header"
  for b in nfllib_demo_main_op1024_60_uint32_t nfllib_demo_main_func1024_60_uint32_t nfllib_demo_main_op8192_124_uint64_t \
           nfllib_demo_main_func8192_124_uint64_t ntt_multi; do
    timeout 600 tests/cpp/_ref/$b > ${O}_demo_$b.log 2>&1; echo "$b rc=$?"; tail -4 ${O}_demo_$b.log
  done
  echo "== Gaussian tests (incl. the large degrees)"
  timeout 900 python -m pytest tests/test_gaussian.py -m gpu -x -q 2>&1 | tail -5
  echo "== memcheck: Gaussian sampler"
  timeout 480 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gaussian.py -m gpu -x -q > ${O}_memcheck_gaussian.log 2>&1; echo "rc=$?"; tail -6 ${O}_memcheck_gaussian.log
  echo "== memcheck: transforms of every size + unit scheduler"
  timeout 540 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sizes or scheduler or edge" > ${O}_memcheck_ntt.log 2>&1; echo "rc=$?"; tail -6 ${O}_memcheck_ntt.log
  echo "== racecheck: transforms of every size + unit scheduler"
  timeout 540 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sizes or scheduler" > ${O}_racecheck_ntt.log 2>&1; echo "rc=$?"; tail -6 ${O}_racecheck_ntt.log
  echo "== racecheck: Gaussian sampler"
  timeout 480 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gaussian.py -m gpu -x -q > ${O}_racecheck_gaussian.log 2>&1; echo "rc=$?"; tail -6 ${O}_racecheck_gaussian.log
} 2>&1 | tee ${O}_call1.log

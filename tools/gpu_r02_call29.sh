#!/bin/bash
# Round 2, GPU call 29 (1 GPU): compute-sanitizer over the kernels changed last (cluster inverse with N^-1 folded into its twiddles; inverse
# launch geometry / table stride refactor): torch-free launch stress (tests/cpp/sched_stress.cpp), cluster shape under memcheck and
# racecheck, every shape under memcheck.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02ac
{
  echo "== cluster shape (u64 N=32768 M=2 batch=200), memcheck"
  timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 tests/cpp/sched_stress 8 32768 2>&1 | tail -4; echo "rc=${PIPESTATUS[0]}"
  echo "== cluster shape, racecheck"
  timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 tests/cpp/sched_stress 4 32768 2>&1 | tail -4; echo "rc=${PIPESTATUS[0]}"
  echo "== every shape, memcheck"
  timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 tests/cpp/sched_stress 8 2>&1 | tail -16; echo "rc=${PIPESTATUS[0]}"
} 2>&1 | tee ${O}_call29.log

#!/bin/bash
# Round 2, GPU call 3: issue ceiling of the butterfly code, slot finish-time trace, launch-geometry variants over all sizes,
# racecheck / memcheck of the torch-free scheduler stress.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02c
{
  echo "== butterfly issue ceiling (tools/ubench/bfly_ceiling.cu)"
  timeout 300 tools/ubench/bfly_ceiling
  echo "== slot finish times, forward N=1024 u64 M=4 batch 4096 (512 threads x 2 CTAs per SM)"
  timeout 300 python tools/trace_slots.py build/variants/trace10/libnflgpu.so
  echo "== launch geometry: tree vs (1024 threads x 1 CTA) vs (512 x 2) vs (256 x 4)"
  for cfg in "64 1024 4 4096" "64 2048 4 2048" "64 4096 4 1024" "64 8192 6 2048" "64 16384 8 256" "64 32768 2 256" "32 1024 8 8192" "32 4096 14 2048" "32 32768 4 512" "16 512 2 16384"; do
    set -- $cfg
    echo "# u$1 N=$2 M=$3 batch=$4"
    for v in nfllib_b200 build/variants/g1024 build/variants/g512 build/variants/g256; do
      timeout 300 python tools/kbench.py --bits $1 --degree $2 --nmoduli $3 --batch $4 --lib $v/libnflgpu.so 2>&1 | tail -1
    done
  done
  echo "== scheduler stress, plain / memcheck / racecheck"
  timeout 300 tests/cpp/sched_stress 300; echo "rc=$?"
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 tests/cpp/sched_stress 60 > ${O}_memcheck_sched.log 2>&1; echo "memcheck rc=$?"; tail -12 ${O}_memcheck_sched.log
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 tests/cpp/sched_stress 40 > ${O}_racecheck_sched.log 2>&1; echo "racecheck rc=$?"; tail -12 ${O}_racecheck_sched.log
} 2>&1 | tee ${O}_call3.log

#!/bin/bash
# Round 2, GPU call 2: the round's new GPU tests, the new bench line (configs / e2e extras), the fixed pipe micro-benchmark,
# occupancy / radix variants of the headline kernel, and the one racecheck run that crashed in call 1.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02b
{
  echo "== new GPU tests"
  timeout 1200 python -m pytest tests/test_round2.py tests/test_reference_programs.py tests/test_gaussian.py -m gpu -x -q -k "not reference_program_passes" 2>&1 | tail -15
  echo "== bench (N=1)"
  timeout 900 python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$?"; cat ${O}_bench.json; tail -5 ${O}_bench.err
  echo "== pipe micro-benchmarks"
  timeout 300 tools/ubench/pipemix
  echo "== variants, N=1024 u64 M=4 batch 4096"
  for v in base10 r80 e3a e3b t1024 base10; do timeout 300 python tools/kbench.py --bits 64 --degree 1024 --nmoduli 4 --batch 4096 --lib build/variants/$v/libnflgpu.so 2>&1 | tail -1; done
  echo "== variants, N=4096 u32 M=14 batch 2048"
  for v in base12 m6_12 m8_12 base12; do timeout 300 python tools/kbench.py --bits 32 --degree 4096 --nmoduli 14 --batch 2048 --lib build/variants/$v/libnflgpu.so 2>&1 | tail -1; done
  echo "== racecheck: unit scheduler test alone"
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scheduler" > ${O}_racecheck_sched.log 2>&1; echo "rc=$?"; tail -25 ${O}_racecheck_sched.log
} 2>&1 | tee ${O}_call2.log

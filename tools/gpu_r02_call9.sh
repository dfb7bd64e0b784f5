#!/bin/bash
# Round 2, GPU call 9 (1 GPU): residue-hopping grid check (kernel table), whole GPU suite, bench, sanitizer pass over the scheduler stress.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02i
{
  echo "== all-config kernel table"
  timeout 900 python tools/kbench_all.py 2>&1 | tee ${O}_kbench_all.txt | cut -c1-130
  echo "== GPU suite"
  timeout 2400 python -m pytest tests -m gpu -x -q > ${O}_pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 ${O}_pytest_gpu.log
  echo "== scheduler stress under racecheck / memcheck (hop + cluster kernels included)"
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 tests/cpp/sched_stress 40 2>&1 | tail -4
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 tests/cpp/sched_stress 24 2>&1 | tail -4
  echo "== bench (N=1)"
  timeout 900 python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$?"; cut -c1-600 ${O}_bench.json; tail -3 ${O}_bench.err
} 2>&1 | tee ${O}_call9.log

#!/bin/bash
# Round 2, GPU call 18 (1 GPU): CUDA-graph capture test, then the whole GPU suite and the sanitizers once more over the tree as it ships
# (128-bit pass 0 for N = 16384 went in after call 15).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02r
{
  echo "== CUDA graph capture"
  timeout 600 python -m pytest tests/test_round2.py -m gpu -x -q -k "cuda_graph" 2>&1 | tail -15
  echo "== GPU suite"
  timeout 2400 python -m pytest tests -m gpu -q > ${O}_pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 ${O}_pytest_gpu.log
  echo "== launch stress: memcheck / racecheck (adds the N = 16384 shape)"
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 tests/cpp/sched_stress 16 2>&1 | tail -13
  timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 tests/cpp/sched_stress 8 2>&1 | tail -13
  echo "== smoke"
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
} 2>&1 | tee ${O}_call18.log

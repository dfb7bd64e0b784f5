#!/bin/bash
# Round 2, GPU call 24 (1 GPU): the tree as it ships, once more after the small-call path: whole GPU suite, launch stress under memcheck / racecheck,
# smoke, bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02w
{
  echo "== GPU suite"
  timeout 2400 python -m pytest tests -m gpu -q > ${O}_pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 ${O}_pytest_gpu.log
  echo "== launch stress: memcheck / racecheck"
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 tests/cpp/sched_stress 16 2>&1 | tail -13
  timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 tests/cpp/sched_stress 8 2>&1 | tail -13
  echo "== smoke"
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
  echo "== bench (N=1)"
  /usr/bin/time -v timeout 900 python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$?"; cut -c1-300 ${O}_bench.json; grep -E "Elapsed|Maximum resident" ${O}_bench.err
  echo "== reference arm"
  timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | cut -c1-500
} 2>&1 | tee ${O}_call24.log

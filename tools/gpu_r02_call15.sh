#!/bin/bash
# Round 2, GPU call 15 (1 GPU): the tree as it ships -- whole GPU suite, sanitizers over the launch stress (pipelined kernels with several units
# per slot, the host ring), kernel table, bench, ncu launch list of the bench command.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02o
{
  echo "== GPU suite"
  timeout 2400 python -m pytest tests -m gpu -x -q > ${O}_pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 ${O}_pytest_gpu.log
  echo "== launch stress: plain / memcheck / racecheck"
  timeout 600 tests/cpp/sched_stress 80 2>&1 | tail -12
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 tests/cpp/sched_stress 24 2>&1 | tail -12
  timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 tests/cpp/sched_stress 8 2>&1 | tail -12
  echo "== all-config kernel table"
  timeout 900 python tools/kbench_all.py 2>&1 | tee ${O}_kbench_all.txt | cut -c1-130
  echo "== bench (N=1)"
  timeout 900 python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$?"; cut -c1-300 ${O}_bench.json; tail -3 ${O}_bench.err
  echo "== ncu launch list of the bench command"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${O}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu > ${O}_ncu_bench.log 2>&1; echo "rc=$?"
  wc -l ${O}_launches.csv
} 2>&1 | tee ${O}_call15.log

#!/bin/bash
# Round 2, GPU call 32 (1 GPU): evidence files of the shipped tree: per-launch times of every kernel on the BASELINE shapes (tools/kbench_all.py)
# and the ncu launch list of the bench command.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02ab
{
  echo "== all-config kernel table"
  timeout 300 python tools/kbench_all.py 2>&1 | tee ${O}_kbench_all.txt | cut -c1-130
  echo "== ncu launch list of the bench command"
  timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${O}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu > ${O}_ncu_bench.log 2>&1; echo "rc=$?"
  wc -l ${O}_launches.csv
} 2>&1 | tee ${O}_call32.log

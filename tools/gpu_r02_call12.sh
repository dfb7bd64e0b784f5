#!/bin/bash
# Round 2, GPU call 12 (1 GPU): ramped chunk schedule of the host pipeline (sweep), whole GPU suite with the swizzled 32-bit N = 4096
# kernels and the ring pipeline, kernel table, ncu capture of the swizzled C4 kernels, bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02l
{
  echo "== e2e sweep: ramp off / on"
  for r in 0 1; do
    echo "-- NFLGPU_HOST_RAMP=$r"
    NFLGPU_HOST_RAMP=$r timeout 600 python tools/e2e_sweep.py 8:4 16:3 16:4 32:3 2>&1
  done | tee ${O}_e2e_sweep.txt
  echo "== GPU suite"
  timeout 2400 python -m pytest tests -m gpu -x -q > ${O}_pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 ${O}_pytest_gpu.log
  echo "== all-config kernel table"
  timeout 900 python tools/kbench_all.py 2>&1 | tee ${O}_kbench_all.txt | cut -c1-130
  echo "== ncu, C4 shape with the swizzled tile"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_ -s 4 -c 2 -f -o /tmp/c4 python tools/kbench.py --bits 32 --degree 4096 --nmoduli 14 --batch 2048 --iters 2 > /tmp/c4.log 2>&1; echo "rc=$?"
  python tools/ncu_summary.py /tmp/c4.ncu-rep > ${O}_ncu_c4.txt 2>&1
  python tools/ncu_stalls.py /tmp/c4.ncu-rep >> ${O}_ncu_c4.txt 2>&1
  grep -E "^==|bank_conflicts|time_duration|issue_active|pipe_lsu|mio_throttle|short_scoreboard" ${O}_ncu_c4.txt | head -20
  echo "== bench (N=1)"
  timeout 900 python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$?"; cut -c1-400 ${O}_bench.json; tail -3 ${O}_bench.err
} 2>&1 | tee ${O}_call12.log

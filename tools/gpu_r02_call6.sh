#!/bin/bash
# Round 2, GPU call 6 (1 GPU): all-config kernel table, bench line, ncu launch list of the bench command, ncu --set full captures
# of the headline kernels and of the C3 / C5 / cluster (N = 2^15) kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02f
{
  echo "== all-config kernel table"
  timeout 900 python tools/kbench_all.py 2>&1 | tee ${O}_kbench_all.txt
  echo "== bench (N=1)"
  timeout 900 python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$?"; cut -c1-1200 ${O}_bench.json; tail -3 ${O}_bench.err
  echo "== ncu launch list of the bench command"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${O}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu > ${O}_ncu_bench.log 2>&1; echo "rc=$?"; tail -2 ${O}_ncu_bench.log | cut -c1-300
  echo "== ncu --set full: C2 forward + inverse"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_ -s 4 -c 2 -f -o ${O}_prof_c2 \
      python tools/kbench.py --iters 2 > ${O}_ncu_c2.log 2>&1; echo "rc=$?"
  echo "== ncu --set full: C3 (N=16384 u64 M=8)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_ -s 4 -c 2 -f -o ${O}_prof_c3 \
      python tools/kbench.py --bits 64 --degree 16384 --nmoduli 8 --batch 256 --iters 2 > ${O}_ncu_c3.log 2>&1; echo "rc=$?"
  echo "== ncu --set full: C5 (N=8192 u64 M=6)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_ -s 4 -c 2 -f -o ${O}_prof_c5 \
      python tools/kbench.py --bits 64 --degree 8192 --nmoduli 6 --batch 512 --iters 2 > ${O}_ncu_c5.log 2>&1; echo "rc=$?"
  echo "== ncu --set full: cluster kernels (N=32768 u64 M=2)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_cluster -s 4 -c 2 -f -o ${O}_prof_n15 \
      python tools/kbench.py --bits 64 --degree 32768 --nmoduli 2 --batch 256 --iters 2 > ${O}_ncu_n15.log 2>&1; echo "rc=$?"
  echo "== ncu --set full: C4 (N=4096 u32 M=14)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_ -s 4 -c 2 -f -o ${O}_prof_c4 \
      python tools/kbench.py --bits 32 --degree 4096 --nmoduli 14 --batch 2048 --iters 2 > ${O}_ncu_c4.log 2>&1; echo "rc=$?"
  ls -la gpurun_out | grep r02f
} 2>&1 | tee ${O}_call6.log

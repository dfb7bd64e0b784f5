#!/bin/bash
# Round 2, GPU call 7 (1 GPU): ncu launch list of the bench command + ncu --set full captures, summarised on the box (the
# reports themselves stay in /tmp except the headline one: gpurun brings back at most 64 MiB).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02g
{
  echo "== ncu launch list of the bench command"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${O}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu > ${O}_ncu_bench.log 2>&1; echo "rc=$?"
  cap() {  # name, kernel regex, kbench args...
    name=$1; rx=$2; shift 2
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 2 -f -o /tmp/${name} python tools/kbench.py "$@" --iters 2 > /tmp/${name}.log 2>&1
    echo "$name rc=$?"
    python tools/ncu_summary.py /tmp/${name}.ncu-rep > ${O}_ncu_${name}.txt 2>&1
    python tools/ncu_stalls.py /tmp/${name}.ncu-rep > ${O}_stalls_${name}.txt 2>&1
  }
  cap c2 ntt_ --bits 64 --degree 1024 --nmoduli 4 --batch 4096
  cp /tmp/c2.ncu-rep ${O}_prof_c2.ncu-rep
  cap c3 ntt_ --bits 64 --degree 16384 --nmoduli 8 --batch 256
  cap c5 ntt_ --bits 64 --degree 8192 --nmoduli 6 --batch 512
  cap n15 ntt_cluster --bits 64 --degree 32768 --nmoduli 2 --batch 256
  cap c4 ntt_ --bits 32 --degree 4096 --nmoduli 14 --batch 2048
  ls -la gpurun_out | grep r02g
} 2>&1 | tee ${O}_call7.log

#!/bin/bash
# Round 2, GPU call 19 (1 GPU): ncu --set full captures of the kernels as they ship (C2, C3, C4, C5, N = 2^15 cluster), summarised on the box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02s
{
  cap() {  # name, kernel regex, kbench args...
    name=$1; rx=$2; shift 2
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 2 -f -o /tmp/${name} python tools/kbench.py "$@" --iters 2 > /tmp/${name}.log 2>&1
    echo "$name rc=$?"
    python tools/ncu_summary.py /tmp/${name}.ncu-rep > ${O}_ncu_${name}.txt 2>&1
    python tools/ncu_stalls.py /tmp/${name}.ncu-rep >> ${O}_ncu_${name}.txt 2>&1
  }
  cap c2 ntt_ --bits 64 --degree 1024 --nmoduli 4 --batch 4096
  cap c3 ntt_ --bits 64 --degree 16384 --nmoduli 8 --batch 256
  cap c4 ntt_ --bits 32 --degree 4096 --nmoduli 14 --batch 2048
  cap c5 ntt_ --bits 64 --degree 8192 --nmoduli 6 --batch 512
  cap n15 ntt_cluster --bits 64 --degree 32768 --nmoduli 2 --batch 256
  grep -h -E "^==|time_duration|dram__bytes|issue_active|inst_executed.sum" ${O}_ncu_c*.txt ${O}_ncu_n15.txt | cut -c1-150
} 2>&1 | tee ${O}_call19.log

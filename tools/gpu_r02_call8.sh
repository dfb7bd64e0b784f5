#!/bin/bash
# Round 2, multi-GPU call (run with gpurun --gpus N): the bench line the driver's scaling run will ask for, and the CUDA-IPC gather test
# with its two processes on two different GPUs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
O=gpurun_out/r02h_n$N
{
  nvidia-smi -L
  echo "== bench --gpus $N (torchrun, the driver's command line)"
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 \
      > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$?"; cat ${O}_bench.json; tail -15 ${O}_bench.err
  echo "== reference arm under torchrun"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 5 --warmup 1 \
      2>/dev/null | cut -c1-400
  echo "== CUDA IPC gather test with the two processes on two GPUs"
  timeout 600 python -m pytest tests/test_round2.py -m gpu -q -k "two_processes or gather" 2>&1 | tail -4
} 2>&1 | tee ${O}_call8.log

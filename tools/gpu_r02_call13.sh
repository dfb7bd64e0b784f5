#!/bin/bash
# Round 2, multi-GPU call (gpurun --gpus N): the bench line the driver's scaling run asks for, at N ranks (C4 residue x batch sharded with the
# peer-memory gather, C5 strong-scaled, end-to-end with every rank's copies at once).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
O=gpurun_out/r02m_n$N
{
  nvidia-smi -L | head -8
  nvidia-smi topo -m 2>/dev/null | head -12
  echo "== bench --gpus $N (torchrun, the driver's command line)"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 \
      > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$?"; cut -c1-3000 ${O}_bench.json; tail -8 ${O}_bench.err
} 2>&1 | tee ${O}_call13.log

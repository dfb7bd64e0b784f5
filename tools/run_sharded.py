"""Multi-GPU demonstration of SURVEY.md section 8e on BASELINE.json configs[3] (N=4096, uint32_t, 14 RNS moduli):
residue x batch sharding (nfllib_b200/sharding.py), one process per GPU, one context per rank over ITS residues only
(nflgpu_ctx_create(first_modulus=...)), no collective on the data path; then one NCCL all_gather to rebuild the full RNS
vectors on every rank (what a CRT lift would need) and a check against the CPU oracle on rank 0.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_sharded.py [--batch B]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

import nfllib_b200 as nb
from nfllib_b200 import sharding as sh


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bits, N, M, batch = 32, 4096, 14, args.batch
    from oracle_lib import Oracle, random_polys  # checker only (rank 0) + seeded input generator
    full = random_polys(bits, N, M, batch, 4242)  # every rank generates the same seeded input, keeps only its shard
    shard = sh.shard_residues(batch, M, world, rank)
    ctx = nb.Context(bits, N, shard.nres, device=local, first_modulus=shard.res0)
    mine = torch.from_numpy(sh.local_view(full, shard).view(np.int32)).cuda()
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        ctx.ntt_fwd(mine.data_ptr(), mine.data_ptr(), shard.npolys, s)
        ctx.ntt_inv(mine.data_ptr(), mine.data_ptr(), shard.npolys, s)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        ctx.ntt_fwd(mine.data_ptr(), mine.data_ptr(), shard.npolys, s)
        ctx.ntt_inv(mine.data_ptr(), mine.data_ptr(), shard.npolys, s)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ctx.ntt_fwd(mine.data_ptr(), mine.data_ptr(), shard.npolys, s)  # leave the forward image in place
    torch.cuda.synchronize()
    g0 = time.perf_counter()
    gathered = sh.gather_residues(mine, shard, batch, M, world)       # the only collective: NCCL all_gather
    torch.cuda.synchronize()
    g1 = time.perf_counter()
    if rank == 0:
        o = Oracle(bits, N, M)
        k = min(batch, 8)
        ok = bool(np.array_equal(gathered[:k].cpu().numpy().view(np.uint32), o.run("fwd", full[:k]))) and \
            bool(np.array_equal(gathered[batch - k:].cpu().numpy().view(np.uint32), o.run("fwd", full[batch - k:])))
        ms = float(t.item())
        print(f"C4 sharded: world={world} shard={shard} fwd+inv {ms:.3f} ms per pass over {batch} polys -> "
              f"{2 * batch / ms / 1e3:.3f} M transforms/s (max over ranks), gather {1e3 * (g1 - g0):.1f} ms, matches oracle: {ok}")
        assert ok
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Host-buffer pipeline tuning sweep (development aid): nflgpu_host_op(fwd) on pinned memory vs chunk size."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import nfllib_b200 as nb
from oracle_lib import random_polys
bits, N, M, batch = 64, 1024, 4, 4096
a = torch.from_numpy(random_polys(bits, N, M, batch, 3).view(np.int64)).pin_memory()
o = torch.empty_like(a).pin_memory()
na, no = a.numpy().view(np.uint64), o.numpy().view(np.uint64)
for mib in (0, 8, 16, 32):
    os.environ["NFLGPU_HOST_ZEROCOPY"] = "1" if mib == 0 else "0"
    os.environ["NFLGPU_HOST_CHUNK_MIB"] = str(max(mib, 1))
    ctx = nb.Context(bits, N, M)
    for _ in range(3):
        ctx.host_op("fwd", na, out=no)
    from oracle_lib import Oracle
    assert np.array_equal(no[:2], Oracle(bits, N, M).run("fwd", na[:2]))
    t0 = time.perf_counter()
    it = 10
    for _ in range(it):
        ctx.host_op("fwd", na, out=no)
    dt = (time.perf_counter() - t0) / it
    print(f"{'zero-copy' if mib == 0 else 'staged   '} chunk {mib:3d} MiB: {dt * 1e3:7.3f} ms per host_op(fwd) of {na.nbytes >> 20} MiB  -> {na.nbytes / dt / 1e9:6.1f} GB/s each way, {batch / dt / 1e6:6.3f} M transforms/s", flush=True)
    ctx.close()

"""Latency of ONE host polynomial through nflgpu_host_op (what poly::ntt_pow_phi() on a host poly costs), per shape and operation,
pageable and pinned, against the oracle (development aid).  NFLGPU_HOST_SMALL_KIB=0 disables the mapped-memory path."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import nfllib_b200 as nb
from oracle_lib import Oracle, random_polys

for bits, N, M in ((64, 1024, 4), (32, 1024, 2), (64, 8192, 2), (16, 128, 1), (64, 16384, 1)):
    ctx, o = nb.Context(bits, N, M), Oracle(bits, N, M)
    a = random_polys(bits, N, M, 1, 5)
    b = random_polys(bits, N, M, 1, 6)
    res = []
    for kind in ("pageable", "pinned"):
        if kind == "pinned":
            view = {16: np.int16, 32: np.int32, 64: np.int64}[bits]
            ta, tb = torch.from_numpy(a.view(view)).pin_memory(), torch.from_numpy(b.view(view)).pin_memory()
            x, y = ta.numpy().view(a.dtype), tb.numpy().view(a.dtype)
            out = torch.empty_like(ta).pin_memory().numpy().view(a.dtype)
        else:
            x, y, out = a.copy(), b.copy(), np.empty_like(a)
        for op, args in (("fwd", (x,)), ("inv", (x,)), ("mul", (x, y)), ("polymul", (x, y))):
            for _ in range(20):
                ctx.host_op(op, *args, out=out)
            assert np.array_equal(out, o.run(op, *args)), (bits, N, M, kind, op)
            lat = []
            for _ in range(200):
                t0 = time.perf_counter()
                ctx.host_op(op, *args, out=out)
                lat.append(time.perf_counter() - t0)
            lat.sort()
            res.append(f"{kind[:4]} {op} {lat[100] * 1e6:5.1f}")
    print(f"u{bits} N={N} M={M} ({a.nbytes >> 10} KiB): " + " | ".join(res) + "  us (median of 200)", flush=True)
    ctx.close()

"""Pageable end-to-end rate against the size of the staging-copy thread pool (development aid).  One step = forward of one 128 MiB
batch + inverse of another on PAGEABLE numpy arrays (the layout of posix_memalign'ed nfl::poly[]), C2 shape, through nflgpu_host_op
(blocking) and nflgpu_host_op_async (K steps queued, one wait).  The pool size is read once per process: one subprocess per value.
usage: python tools/pageable_sweep.py [--threads 2,4,8,12,16]"""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child():
    import numpy as np
    from oracle_lib import Oracle, random_polys
    import nfllib_b200 as nb
    bits, N, M, batch, steps = 64, 1024, 4, 4096, 6
    ctx = nb.Context(bits, N, M, device=0)
    o = Oracle(bits, N, M)
    pA = random_polys(bits, N, M, batch, 11)
    pD = random_polys(bits, N, M, batch, 12)
    pB, pC = np.empty_like(pA), np.empty_like(pD)

    def blocking():
        ctx.host_op("fwd", pA, out=pB)
        ctx.host_op("inv", pD, out=pC)

    def queued():
        ctx.host_op("fwd", pA, out=pB, wait=False)
        ctx.host_op("inv", pD, out=pC, wait=False)

    res = {}
    for name, fn in (("blocking", blocking), ("async", queued)):
        for _ in range(2):
            fn()
        ctx.host_sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        ctx.host_sync()
        res[name] = 2.0 * batch * steps / (time.perf_counter() - t0)
    ok = bool(np.array_equal(pB[-2:], o.run("fwd", pA[-2:]))) and bool(np.array_equal(pC[:2], o.run("inv", pD[:2])))
    print(f"NFLGPU_HOST_COPY_THREADS={os.environ.get('NFLGPU_HOST_COPY_THREADS', 'default'):>7s}: blocking {res['blocking'] / 1e6:.3f} M transforms/s, "
          f"async {res['async'] / 1e6:.3f} M transforms/s, {'OK' if ok else 'MISMATCH'}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child()
        sys.exit(0)
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", default="default,2,4,8,12,16")
    a = ap.parse_args()
    print(f"host threads: {os.cpu_count()} (affinity {len(os.sched_getaffinity(0))})", flush=True)
    for t in a.threads.split(","):
        env = dict(os.environ)
        if t == "default":
            env.pop("NFLGPU_HOST_COPY_THREADS", None)
        else:
            env["NFLGPU_HOST_COPY_THREADS"] = t
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, timeout=300)

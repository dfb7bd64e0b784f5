"""Where the mapped-memory path of small host calls stops paying (development aid): nflgpu_host_op(fwd) on 1 .. 256 C2 polynomials
(32 KiB each), pinned and pageable, with NFLGPU_HOST_SMALL_KIB = 0 (copy engines) and = 16384 (kernel reads / writes host memory)."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def child():
    import numpy as np, torch
    import nfllib_b200 as nb
    from oracle_lib import Oracle, random_polys
    bits, N, M = 64, 1024, 4
    ctx, o = nb.Context(bits, N, M), Oracle(bits, N, M)
    out = []
    for batch in (1, 2, 4, 8, 16, 32, 64, 128, 256):
        a = random_polys(bits, N, M, batch, 11)
        ta = torch.from_numpy(a.view(np.int64)).pin_memory()
        pa, po = ta.numpy().view(np.uint64), torch.empty_like(ta).pin_memory().numpy().view(np.uint64)
        ga, go = a.copy(), np.empty_like(a)
        row = []
        for x, y in ((pa, po), (ga, go)):
            for _ in range(10):
                ctx.host_op("fwd", x, out=y)
            assert np.array_equal(y[-1:], o.run("fwd", x[-1:]))
            lat = []
            for _ in range(100):
                t0 = time.perf_counter()
                ctx.host_op("fwd", x, out=y)
                lat.append(time.perf_counter() - t0)
            lat.sort()
            row.append(lat[50] * 1e6)
        out.append(f"{batch * 32:5d} KiB: {row[0]:7.1f} / {row[1]:7.1f}")
    print(f"NFLGPU_HOST_SMALL_KIB={os.environ['NFLGPU_HOST_SMALL_KIB']:>5s}  pinned / pageable us | " + " | ".join(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child()
    else:
        for k in ("0", "16384", "0", "16384"):
            subprocess.call([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, NFLGPU_HOST_SMALL_KIB=k))

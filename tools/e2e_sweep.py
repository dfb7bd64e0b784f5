"""Host-buffer pipeline tuning sweep (development aid): the bench's e2e step (fwd of one 128 MiB batch + inv of another, pinned
host memory) against chunk size and ring depth, in three call patterns:
  sync   nflgpu_host_op(fwd); nflgpu_host_op(inv)                     (every call waits for its own last download)
  step   nflgpu_host_op_async(fwd); nflgpu_host_op_async(inv); sync  (one wait per step)
  stream the async pair for every step, one nflgpu_host_sync at the end
The knobs are read once per process, so every configuration runs in a child process."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def child():
    import numpy as np, torch
    import nfllib_b200 as nb
    import nfllib_b200.capi as capi
    from oracle_lib import Oracle, random_polys
    check = True
    if os.environ.get("E2E_LIB"):  # an experiment build (tools/variants.sh); "nokernel" builds skip the result check
        capi.lib_path = lambda: os.path.abspath(os.environ["E2E_LIB"])
        check = "nokernel" not in os.environ["E2E_LIB"]
    bits, N, M, batch = 64, 1024, 4, 4096
    pin = lambda x: torch.from_numpy(x.view(np.int64)).pin_memory()
    a, d = pin(random_polys(bits, N, M, batch, 3)), pin(random_polys(bits, N, M, batch, 4))
    b, c = torch.empty_like(a).pin_memory(), torch.empty_like(a).pin_memory()
    na, nb_, nc, nd = (t.numpy().view(np.uint64) for t in (a, b, c, d))
    ctx = nb.Context(bits, N, M)
    o = Oracle(bits, N, M)
    it = 10

    def sync_step():
        ctx.host_op("fwd", na, out=nb_)
        ctx.host_op("inv", nd, out=nc)

    def async_pair():
        ctx.host_op("fwd", na, out=nb_, wait=False)
        ctx.host_op("inv", nd, out=nc, wait=False)

    def step_step():
        async_pair()
        ctx.host_sync()

    res = {}
    for name, fn, tail in (("sync", sync_step, None), ("step", step_step, None), ("stream", async_pair, ctx.host_sync)):
        for _ in range(3):
            fn()
        ctx.host_sync()
        t0 = time.perf_counter()
        for _ in range(it):
            fn()
        if tail:
            tail()
        res[name] = (time.perf_counter() - t0) / it
        if check:
            assert np.array_equal(nb_[:2], o.run("fwd", na[:2])) and np.array_equal(nb_[-1:], o.run("fwd", na[-1:]))
            assert np.array_equal(nc[:2], o.run("inv", nd[:2])) and np.array_equal(nc[-1:], o.run("inv", nd[-1:]))
    one_in, one_out = np.array(na[:1]), np.empty_like(na[:1])
    for _ in range(20):
        ctx.host_op("fwd", one_in, out=one_out)
    if check:
        assert np.array_equal(one_out, o.run("fwd", one_in))
    lat = []
    for _ in range(200):
        t0 = time.perf_counter()
        ctx.host_op("fwd", one_in, out=one_out)
        lat.append(time.perf_counter() - t0)
    lat.sort()
    zc = os.environ.get("NFLGPU_HOST_ZEROCOPY") == "1"
    tag = ("" if check else "NO KERNEL ") + "zero-copy          " if zc else ("" if check else "NO KERNEL ") + f"chunk {int(os.environ['NFLGPU_HOST_CHUNK_MIB']):3d} MiB ring {int(os.environ['NFLGPU_HOST_RING']):2d}"
    print(tag + " | " + " | ".join(f"{k} {v * 1e3:6.3f} ms/step {2 * batch / v / 1e6:6.3f} M tr/s" for k, v in res.items()) +
          f" | one poly {lat[100] * 1e6:5.1f} us", flush=True)
    ctx.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    grid = [(c, r) for c in (1, 2, 4, 8, 16) for r in (4, 8, 16)] if len(sys.argv) < 2 else [tuple(map(int, x.split(":"))) for x in sys.argv[1:]]
    for chunk, ring in grid:
        env = dict(os.environ, NFLGPU_HOST_CHUNK_MIB=str(chunk), NFLGPU_HOST_RING=str(ring), NFLGPU_HOST_ZEROCOPY="0")
        subprocess.call([sys.executable, os.path.abspath(__file__), "child"], env=env)

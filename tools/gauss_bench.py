"""Times nflgpu_gaussian_sample (development aid): FastGaussianNoise<uint8_t, uint64_t, 2>(20, 128, 2^14) draws for the C2 shape."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nfllib_b200.capi as capi

ctx = capi.Context(64, 1024, 4)
g = capi.Gaussian(ctx, 20.0, 128, 1 << 14, 0.0, 1, 2)
key = bytes(range(1, 33))
for batch in (256, 4096):
    p = ctx.alloc(batch)
    g.sample(p, batch, key, 0)
    t0 = time.perf_counter()
    reps = 5
    for r in range(reps):
        used = g.sample(p, batch, key, 1000 * r)
    dt = (time.perf_counter() - t0) / reps
    print(f"gaussian batch {batch}: {dt * 1e3:.2f} ms per call (synchronous), {batch / dt / 1e6:.2f} M polys/s, {used} nonces")
    ctx.free(p)

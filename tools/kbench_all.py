"""All-config kernel timing table (development aid): BASELINE.json configs C2..C5 x {fwd, inv, mul, add, polymul},
device-resident, L2-cold rotation, CUDA events; achieved GB/s uses the algorithmic bytes of SURVEY.md section 8d."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import nfllib_b200 as nb
import nfllib_b200.capi as capi
from oracle_lib import random_polys

if os.environ.get("NFLGPU_LIB"):  # experiment builds (tools/variants.sh): time another libnflgpu.so
    capi.lib_path = lambda: os.path.abspath(os.environ["NFLGPU_LIB"])

CONFIGS = [("C2", 64, 1024, 4, 4096), ("C3", 64, 16384, 8, 1024), ("C4", 32, 4096, 14, 8192), ("C5", 64, 8192, 6, 2048),
           ("u64_2k", 64, 2048, 4, 2048), ("u64_4k", 64, 4096, 4, 1024), ("u64_32k", 64, 32768, 2, 256), ("u32_1k", 32, 1024, 8, 8192), ("u32_32k", 32, 32768, 4, 512)]
if len(sys.argv) > 1:
    CONFIGS = [c for c in CONFIGS if c[0] in sys.argv[1].split(",")]
IT = {16: np.int16, 32: np.int32, 64: np.int64}
for name, bits, N, M, batch in CONFIGS:
    ctx = nb.Context(bits, N, M)
    one = random_polys(bits, N, M, min(batch, 256), 5)
    host = np.concatenate([one] * (batch // one.shape[0]))
    nbytes = host.nbytes
    R = max(2, min(3, int(3e9 // (3 * nbytes))))
    a = [torch.from_numpy(host.view(IT[bits])).cuda() for _ in range(R)]
    b = [torch.from_numpy(host.view(IT[bits])).cuda() for _ in range(R)]
    d = [torch.empty_like(a[0]) for _ in range(R)]
    s = torch.cuda.current_stream().cuda_stream
    ops = {"fwd": (lambda i: ctx.ntt_fwd(d[i].data_ptr(), a[i].data_ptr(), batch, s), 2),
           "inv": (lambda i: ctx.ntt_inv(d[i].data_ptr(), a[i].data_ptr(), batch, s), 2),
           "mul": (lambda i: ctx.mul(d[i].data_ptr(), a[i].data_ptr(), b[i].data_ptr(), batch, s), 3),
           "add": (lambda i: ctx.add(d[i].data_ptr(), a[i].data_ptr(), b[i].data_ptr(), batch, s), 3),
           "muladd": (lambda i: ctx.muladd(d[i].data_ptr(), a[i].data_ptr(), b[i].data_ptr(), a[(i + 1) % R].data_ptr(), batch, s), 4),
           "eval(a+b*c)": (lambda i: ctx.eval(d[i].data_ptr(), [a[i].data_ptr(), b[i].data_ptr(), a[(i + 1) % R].data_ptr()], [0, 1, 2, 0x12, 0x10], batch, s), 4),
           "polymul": (lambda i: ctx.polymul(d[i].data_ptr(), a[i].data_ptr(), b[i].data_ptr(), batch, s), 3)}
    line = f"{name:8s} u{bits} N={N:5d} M={M:2d} batch={batch:5d} ({nbytes >> 20:5d} MiB) "
    for op, (fn, passes) in ops.items():
        for i in range(3):
            fn(i % R)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for i in range(iters):
            fn(i % R)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        line += f"| {op} {us:8.1f}us {passes * nbytes / us / 1e3:6.0f}GB/s "
    print(line, flush=True)
    ctx.close()
    del a, b, d
    torch.cuda.empty_cache()

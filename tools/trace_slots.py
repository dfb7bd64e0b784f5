"""Development aid: when do the unit slots of the forward kernel finish?  Needs a library built with -DNFLGPU_TRACE
(tools/variants.sh 10 trace10 "-DNFLGPU_TRACE").  usage: python tools/trace_slots.py build/variants/trace10/libnflgpu.so"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import nfllib_b200.capi as capi
from oracle_lib import random_polys

capi.lib_path = lambda: os.path.abspath(sys.argv[1])
bits, N, M, batch = 64, 1024, 4, 4096
ctx = capi.Context(bits, N, M)
host = random_polys(bits, N, M, batch, 77)
src = [torch.from_numpy(host.view(np.int64)).cuda() for _ in range(3)]
dst = [torch.empty_like(src[0]) for _ in range(3)]
L = capi.lib()
L.nflgpu_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
s = torch.cuda.current_stream().cuda_stream
for i in range(5):
    ctx.ntt_fwd(dst[i % 3].data_ptr(), src[i % 3].data_ptr(), batch, s)
torch.cuda.synchronize()
for rep in range(3):
    L.nflgpu_debug_trace(None, 0, 1)
    torch.cuda.synchronize()
    ctx.ntt_fwd(dst[rep].data_ptr(), src[rep].data_ptr(), batch, s)
    torch.cuda.synchronize()
    buf = np.zeros(1 + 8192, np.uint64)
    L.nflgpu_debug_trace(buf.ctypes.data, buf.size, 0)
    t0 = int(buf[0])
    ends = np.sort((buf[1:][buf[1:] > 0].astype(np.int64) - t0) / 1e3)
    q = lambda f: ends[min(len(ends) - 1, int(f * len(ends)))]
    print(f"slots {len(ends)}: finish time after the first CTA start, us: min {ends[0]:.1f} p10 {q(.1):.1f} p25 {q(.25):.1f} median {q(.5):.1f} "
          f"p75 {q(.75):.1f} p90 {q(.9):.1f} p99 {q(.99):.1f} max {ends[-1]:.1f};  mean/max = {ends.mean() / ends[-1]:.3f}")

"""Kernel micro-benchmark for experiment variants (development aid; bench.py is the contract benchmark).
usage: python tools/kbench.py [--lib path/to/libnflgpu.so ...] [--bits 64 --degree 1024 --nmoduli 4 --batch 4096]
Times nflgpu_ntt_fwd / nflgpu_ntt_inv on device-resident, L2-cold operands with CUDA events; checks a slice vs the oracle."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", action="append", default=[])
    ap.add_argument("--bits", type=int, default=64)
    ap.add_argument("--degree", type=int, default=1024)
    ap.add_argument("--nmoduli", type=int, default=4)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--ops", default="fwd,inv")
    a = ap.parse_args()
    from oracle_lib import Oracle, random_polys
    import nfllib_b200.capi as capi
    o = Oracle(a.bits, a.degree, a.nmoduli)
    host = random_polys(a.bits, a.degree, a.nmoduli, a.batch, 77)
    exp_f = o.run("fwd", host[:2])
    exp_i = o.run("inv", host[:2])
    bytes_ = 2 * host.nbytes
    libs = a.lib or [capi.lib_path()]
    for path in libs:
        capi._lib = None
        capi.lib_path = (lambda p: (lambda: p))(os.path.abspath(path))
        ctx = capi.Context(a.bits, a.degree, a.nmoduli)
        R = 3
        src = [torch.from_numpy(host.view({16: np.int16, 32: np.int32, 64: np.int64}[a.bits])).cuda() for _ in range(R)]
        dst = [torch.empty_like(src[0]) for _ in range(R)]
        s = torch.cuda.current_stream().cuda_stream
        res = {}
        for op in a.ops.split(","):
            fn = {"fwd": ctx.ntt_fwd, "inv": ctx.ntt_inv}[op]
            for i in range(3):
                fn(dst[i % R].data_ptr(), src[i % R].data_ptr(), a.batch, s)
            torch.cuda.synchronize()
            got = dst[0][:2].cpu().numpy().view(host.dtype)
            ok = np.array_equal(got, exp_f if op == "fwd" else exp_i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.iters):
                fn(dst[i % R].data_ptr(), src[i % R].data_ptr(), a.batch, s)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / a.iters * 1e3
            res[op] = (us, bytes_ / us / 1e3, ok)
        print(f"{os.path.relpath(path, ROOT):60s} " + "  ".join(f"{k}: {v[0]:8.1f} us {v[1]:7.0f} GB/s {'OK' if v[2] else 'MISMATCH'}" for k, v in res.items()), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()

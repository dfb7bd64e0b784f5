"""Whole-batch parity of an experiment build (development aid): forward and inverse transforms of every polynomial of several batch
sizes against the oracle, through the variant library given with --lib.
usage: python tools/check_variant.py --lib build/variants/x/libnflgpu.so [--bits 64 --degree 1024 --nmoduli 4] [--batches 1,37,4096,4099]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", required=True)
    ap.add_argument("--bits", type=int, default=64)
    ap.add_argument("--degree", type=int, default=1024)
    ap.add_argument("--nmoduli", type=int, default=4)
    ap.add_argument("--batches", default="1,37,4096,4099")
    a = ap.parse_args()
    from oracle_lib import Oracle, random_polys
    import nfllib_b200.capi as capi
    capi._lib = None
    capi.lib_path = lambda: os.path.abspath(a.lib)
    ctx = capi.Context(a.bits, a.degree, a.nmoduli)
    o = Oracle(a.bits, a.degree, a.nmoduli)
    ok = True
    for b in [int(x) for x in a.batches.split(",")]:
        x = random_polys(a.bits, a.degree, a.nmoduli, b, 4242 + b)
        for rep in range(3):  # several launches: the walk's counters must come back to their initial state every time
            f = ctx.run_device("ntt_fwd", x)
            i = ctx.run_device("ntt_inv", x)
            good = np.array_equal(f, o.run("fwd", x)) and np.array_equal(i, o.run("inv", x))
            ok = ok and good
        print(f"batch {b}: {'OK' if good else 'MISMATCH'}", flush=True)
    ctx.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

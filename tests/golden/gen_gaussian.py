"""Regenerates tests/golden/gaussian.npz from the UNMODIFIED reference (oracle/_ref/libnflref.so, built with the MPFR/GMP
runtimes of the image): for a few FastGaussianNoise<in_class, T, lu_depth>(sigma, security, samples, center) objects, the
reference's own barrier table and private parameters, and poly::set(gaussian(&prng, amplifier)) draws with the harness's fixed
Salsa20 key (first nonce, nonces consumed, the polynomials or their sha256).
    python tests/golden/gen_gaussian.py"""
import hashlib
import json
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle_lib import Ref, GOLDEN  # noqa: E402

# name: (sigma, security, samples, center, in_bytes, lu_depth, limb bits, degree, nmoduli, batch, amplifier, keep the draws?)
CASES = {
    "demo_u64": (20.0, 128, 1 << 14, 0.0, 1, 2, 64, 1024, 4, 12, 1, False),      # tests/nfllib_demo_main_op.cpp:141
    "prng_demo_u64_small": (3.19, 128, 1 << 19, 5.25, 1, 2, 64, 64, 3, 8, 2, True),  # tests/prng_demo_main.cpp:10
    "depth1_u16": (3.19, 128, 1 << 10, 0.0, 1, 1, 16, 512, 2, 12, 2, False),
    "words16_u32": (300.0, 128, 1 << 10, 0.0, 2, 1, 32, 4096, 1, 6, 1, False),   # tests/prng_demo_main.cpp:9 (commented shape)
}


def main():
    out, meta = {}, {}
    for name, (sigma, sec, samples, center, ib, depth, bits, N, M, batch, amp, keep) in CASES.items():
        h, t = Ref.gaussian_table(sigma, sec, samples, center, ib, depth, bits)
        first, used, polys = Ref(bits, N, M).gaussian(h, batch, amp)
        out[name + "_barriers"] = t.barriers
        if keep:
            out[name + "_draws"] = polys
        meta[name] = {"sigma": sigma, "security": sec, "samples": samples, "center": center, "in_bytes": ib, "lu_depth": depth, "bits": bits,
                      "N": N, "M": M, "batch": batch, "amplifier": amp, "first_nonce": first, "nonces_used": used,
                      "rounded_center": int(t.rounded_center), "params": t.params,
                      "sha256": hashlib.sha256(np.ascontiguousarray(polys).tobytes()).hexdigest()}
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLDEN, "gaussian.npz"), **out)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()

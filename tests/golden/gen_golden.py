"""Regenerates tests/golden/* from the UNMODIFIED reference compiled into oracle/_ref/libnflref.so.
Run in the build container (needs /root/reference to have been compiled by `make -C oracle ref`):
    python tests/golden/gen_golden.py
Outputs:
  params.json        first 16 entries (2 for uint16_t) of params<T>::{P,Pn,primitive_roots,invkMaxPolyDegree}
  kat_<cfg>.npz      small known-answer sets: inputs + reference outputs of every op on the hot path
  hashes.json        sha256 of reference outputs on seeded inputs at the BASELINE.json configurations
"""
import hashlib
import json
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle_lib import Ref, GOLDEN, DTYPES, random_polys  # noqa: E402


def main():
    params = {bits: Ref.params(bits, 16) for bits in (16, 32, 64)}
    with open(os.path.join(GOLDEN, "params.json"), "w") as f:
        json.dump(params, f, indent=0)

    # known-answer sets (small enough to commit)
    kat_cfgs = [(64, 8, 1), (64, 64, 3), (64, 1024, 1), (64, 1024, 4), (32, 8, 2), (32, 1024, 2), (16, 128, 2), (16, 512, 2),
                (64, 2048, 1), (32, 4096, 1)]
    for bits, N, M in kat_cfgs:
        batch = 3 if N * M <= 4096 else 2
        P = params[bits]["P"]
        a = random_polys(bits, N, M, batch, 1000 + N + M, P)
        b = random_polys(bits, N, M, batch, 2000 + N + M, P)
        # edge polys: all zero, all p-1, delta_0, X (delta_1), delta_{N-1}
        e = np.zeros((5, M, N), DTYPES[bits])
        for cm in range(M):
            e[1, cm, :] = P[cm] - 1
        e[2, :, 0] = 1
        e[3, :, 1] = 1
        e[4, :, N - 1] = 1
        a = np.concatenate([a, e])
        b = np.concatenate([b, e[::-1]])
        r = Ref(bits, N, M)
        fa, fb = r.run("fwd", a), r.run("fwd", b)
        bs = r.run("compute_shoup", b)
        np.savez_compressed(os.path.join(GOLDEN, f"kat_u{bits}_n{N}_m{M}.npz"), a=a, b=b, fwd_a=fa, fwd_b=fb,
                            inv_a=r.run("inv", a), mul=r.run("mul", a, b), add=r.run("add", a, b), sub=r.run("sub", a, b),
                            shoup_b=bs, mul_shoup=r.run("mul_shoup", a, b, bs), polymul=r.run("polymul", a, b),
                            muladd=r.run("muladd", a, b, fa), raw_ntt=r.run("raw_ntt", a), raw_intt=r.run("raw_intt", a))

    # the reference's own uniform sampler, made reproducible by the harness's fixed Salsa20 key (oracle/ref_harness.cpp)
    n0, draws = Ref(64, 1024, 4).uniform(2)
    np.savez_compressed(os.path.join(GOLDEN, "uniform_u64_n1024_m4.npz"), draws=draws, key=np.frombuffer(Ref.FIXED_KEY, dtype=np.uint8),
                        first_nonce=np.uint64(n0))

    # CRT lift by the reference's GMP code (gmp.hpp:183-219)
    lp = random_polys(64, 64, 3, 3, 808, params[64]["P"])
    np.savez_compressed(os.path.join(GOLDEN, "lift_u64_n64_m3.npz"), polys=lp, words=Ref(64, 64, 3).lift(lp, 3))

    # hashes at the BASELINE.json configurations (batch kept small; inputs are seeded, see random_polys)
    hashes = {}
    for name, bits, N, M, batch in [("C1", 64, 1024, 1, 4), ("C2", 64, 1024, 4, 16), ("C3", 64, 16384, 8, 2), ("C4", 32, 4096, 14, 4),
                                    ("C5", 64, 8192, 6, 2)]:
        P = params[bits]["P"]
        a = random_polys(bits, N, M, batch, 7 + N, P)
        b = random_polys(bits, N, M, batch, 11 + N, P)
        r = Ref(bits, N, M)
        h = lambda x: hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()
        fa = r.run("fwd", a)
        hashes[name] = {"bits": bits, "N": N, "M": M, "batch": batch, "seed_a": 7 + N, "seed_b": 11 + N, "in_a": h(a), "fwd_a": h(fa),
                        "inv_fwd_a": h(r.run("inv", fa)), "mul": h(r.run("mul", a, b)), "add": h(r.run("add", a, b)),
                        "sub": h(r.run("sub", a, b)), "polymul": h(r.run("polymul", a, b))}
        assert hashes[name]["inv_fwd_a"] == hashes[name]["in_a"]
    with open(os.path.join(GOLDEN, "hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1)
    print("golden fixtures written to", GOLDEN)


if __name__ == "__main__":
    main()

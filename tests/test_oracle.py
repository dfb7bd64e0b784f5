"""Pins the CPU oracle (oracle/nfl_oracle.c):
  * against committed fixtures that were produced by the unmodified reference (tests/golden/gen_golden.py),
  * against the closed-form definition out[j] = a(phi^(2*bitrev(j)+1)),
  * and, when oracle/_ref/libnflref.so is present, bit-for-bit against the reference itself on every size.
Mirrors the reference's own checks: tests/nfl_{add,sub,mul}.cpp (op vs naive formula), tests/poly_p.cpp:52-58
(NTT round trip) -- but always with full-array equality, never the reference's any-equal operator==."""
import glob
import hashlib
import json
import os
import re

import numpy as np
import pytest

from oracle_lib import GOLDEN, DTYPES, Oracle, Ref, have_ref, golden_params, random_polys, crt_lift, crt_unlift, lift_words_per_coeff

KATS = sorted(glob.glob(os.path.join(GOLDEN, "kat_*.npz")))


def _cfg(path):
    m = re.search(r"kat_u(\d+)_n(\d+)_m(\d+)", path)
    return tuple(int(x) for x in m.groups())


@pytest.mark.parametrize("path", KATS, ids=[os.path.basename(p) for p in KATS])
def test_oracle_matches_reference_fixtures(path):
    bits, N, M = _cfg(path)
    k = np.load(path)
    o = Oracle(bits, N, M)
    a, b = k["a"], k["b"]
    assert np.array_equal(o.run("fwd", a), k["fwd_a"])
    assert np.array_equal(o.run("fwd", b), k["fwd_b"])
    assert np.array_equal(o.run("inv", a), k["inv_a"])
    assert np.array_equal(o.run("raw_ntt", a), k["raw_ntt"])
    assert np.array_equal(o.run("raw_intt", a), k["raw_intt"])
    assert np.array_equal(o.run("mul", a, b), k["mul"])
    assert np.array_equal(o.run("add", a, b), k["add"])
    assert np.array_equal(o.run("sub", a, b), k["sub"])
    assert np.array_equal(o.run("compute_shoup", b), k["shoup_b"])
    assert np.array_equal(o.run("mul_shoup", a, b, k["shoup_b"]), k["mul_shoup"])
    assert np.array_equal(o.run("polymul", a, b), k["polymul"])
    assert np.array_equal(o.run("muladd", a, b, k["fwd_a"]), k["muladd"])
    # round trip (tests/poly_p.cpp:52-58 pattern, but a strong comparison)
    assert np.array_equal(o.run("inv", k["fwd_a"]), a)


@pytest.mark.parametrize("bits,N,M", [(64, 2, 1), (64, 4, 1), (64, 8, 2), (64, 64, 2), (32, 8, 2), (32, 128, 3), (16, 16, 1), (16, 64, 2)])
def test_oracle_matches_closed_form(bits, N, M):
    if N == 2:
        pytest.skip("reference's degree==2 path returns lazily reduced values (core.hpp:468-483)")
    o = Oracle(bits, N, M)
    a = random_polys(bits, N, M, 3, 42 + N)
    assert np.array_equal(o.run("fwd", a), o.spec_fwd(a))


def test_oracle_negacyclic_shift():
    """inv(fwd(a) * fwd(X)) = X*a mod (X^N + 1): coefficients shift up by one, the wrapped one is negated."""
    bits, N, M = 64, 256, 2
    o = Oracle(bits, N, M)
    a = random_polys(bits, N, M, 2, 5)
    x = np.zeros_like(a)
    x[:, :, 1] = 1
    got = o.run("polymul", a, x)
    P = golden_params(bits)["P"]
    exp = np.roll(a, 1, axis=2)
    for cm in range(M):
        exp[:, cm, 0] = (np.uint64(P[cm]) - a[:, cm, N - 1]) % np.uint64(P[cm])
    assert np.array_equal(got, exp)


def test_oracle_hashes_at_baseline_configs():
    """sha256 of reference outputs on seeded inputs at the BASELINE.json shapes (tests/golden/hashes.json)."""
    with open(os.path.join(GOLDEN, "hashes.json")) as f:
        hs = json.load(f)
    h = lambda x: hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()
    for name, e in hs.items():
        if e["N"] * e["M"] * e["batch"] > 1 << 17:
            continue  # keep the CPU suite quick; the GPU parity tests cover every entry
        o = Oracle(e["bits"], e["N"], e["M"])
        a = random_polys(e["bits"], e["N"], e["M"], e["batch"], e["seed_a"])
        b = random_polys(e["bits"], e["N"], e["M"], e["batch"], e["seed_b"])
        assert h(a) == e["in_a"]
        assert h(o.run("fwd", a)) == e["fwd_a"], name
        assert h(o.run("mul", a, b)) == e["mul"], name
        assert h(o.run("polymul", a, b)) == e["polymul"], name


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libnflref.so not built (needs /root/reference)")
@pytest.mark.parametrize("bits", [16, 32, 64])
def test_oracle_matches_live_reference_all_sizes(bits):
    lo = {16: 4, 32: 3, 64: 2}[bits]  # sizeof(poly) is padded to 32 B below these (poly.hpp:88)
    hi = {16: 9, 32: 15, 64: 15}[bits]
    for ln in range(lo, hi + 1):
        N = 1 << ln
        r = Ref(bits, N, 1)
        if not r.supported():
            continue
        o = Oracle(bits, N, 1)
        batch = 2 if N >= 8192 else 4
        a = random_polys(bits, N, 1, batch, 100 + ln)
        b = random_polys(bits, N, 1, batch, 200 + ln)
        for op in ("fwd", "inv", "raw_ntt", "raw_intt"):
            assert np.array_equal(o.run(op, a), r.run(op, a)), (bits, N, op)
        for op in ("mul", "add", "sub"):
            assert np.array_equal(o.run(op, a, b), r.run(op, a, b)), (bits, N, op)
        bs = r.run("compute_shoup", b)
        assert np.array_equal(o.run("compute_shoup", b), bs)
        assert np.array_equal(o.run("mul_shoup", a, b, bs), r.run("mul_shoup", a, b, bs))


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libnflref.so not built (needs /root/reference)")
def test_golden_params_match_live_reference():
    for bits in (16, 32, 64):
        g = golden_params(bits)
        live = Ref.params(bits, 16)
        assert g == live


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libnflref.so not built (needs /root/reference)")
def test_oracle_uniform_sampler_matches_live_reference():
    """poly::set(uniform) of the reference (its own Salsa20 assembly, keyed with the harness's fixed key) vs the oracle's
    Salsa20 + mask/subtract restatement (core.hpp:150-187, fastrandombytes.cpp:21-34)."""
    for bits, N, M in ((64, 1024, 4), (64, 64, 3), (32, 4096, 1), (32, 8, 2), (16, 512, 2), (16, 16, 1)):
        n0, ref = Ref(bits, N, M).uniform(3)
        assert np.array_equal(Oracle(bits, N, M).uniform(3, Ref.FIXED_KEY, n0), ref), (bits, N, M)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libnflref.so not built (needs /root/reference)")
def test_oracle_bounded_and_zo_samplers_match_live_reference():
    for bits, N, M in ((64, 1024, 4), (32, 8, 2), (16, 512, 2)):
        r, o = Ref(bits, N, M), Oracle(bits, N, M)
        for ub, amp in ((1, 1), (2, 1), (5, 3), (1 << 10, 1), (1000, 7)):
            n0, ref = r.sample("non_uniform", 3, ub, amp)
            assert np.array_equal(o.non_uniform(3, ub, amp, Ref.FIXED_KEY, n0), ref), (bits, N, M, ub, amp)
        for rho in (0, 1, 0x7F, 0xFF):
            n0, ref = r.sample("zo", 3, rho)
            assert np.array_equal(o.zo(3, rho, Ref.FIXED_KEY, n0), ref), (bits, N, M, rho)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libnflref.so not built (needs /root/reference)")
def test_oracle_hwt_sampler_matches_live_reference():
    for bits, N, M in ((64, 1024, 4), (32, 8, 2), (16, 512, 2), (64, 64, 3)):
        r, o = Ref(bits, N, M), Oracle(bits, N, M)
        for hwt in (1, 3, N // 4, N - 1, N):
            n0, ref = r.sample("hwt", 2, hwt)
            mine, calls = o.hwt(2, hwt, Ref.FIXED_KEY, n0)
            assert np.array_equal(mine, ref), (bits, N, M, hwt)
            assert calls == 2 * ((N - hwt + hwt - 1) // hwt + 1)
            assert all(np.count_nonzero(mine[b, 0]) == hwt for b in range(2))


def test_oracle_uniform_sampler_fixture():
    k = np.load(os.path.join(GOLDEN, "uniform_u64_n1024_m4.npz"))
    o = Oracle(64, 1024, 4)
    assert np.array_equal(o.uniform(k["draws"].shape[0], bytes(k["key"]), int(k["first_nonce"])), k["draws"])


def _lift_cases(bits, N, M):
    P = golden_params(bits)["P"][:M]
    a = random_polys(bits, N, M, 2, 55)
    e = np.zeros((3, M, N), DTYPES[bits])
    for cm in range(M):
        e[1, cm, :] = P[cm] - 1          # lifts to Q - 1
    e[2, 0, :] = 1                       # only one residue non-zero
    return P, np.concatenate([a, e])


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libnflref.so not built (needs /root/reference)")
def test_crt_lift_restatement_matches_live_reference():
    """poly2mpz / mpz2poly of the reference (its GMP code, gmp.hpp:183-219, linked against the image's libgmp runtime) vs the
    big-integer restatement in oracle_lib."""
    for bits, N, M in ((64, 1024, 4), (64, 64, 3), (32, 1024, 2), (32, 4096, 14), (16, 512, 2)):
        P, a = _lift_cases(bits, N, M)
        W = lift_words_per_coeff(P)
        r = Ref(bits, N, M)
        ref = r.lift(a, W)
        assert np.array_equal(crt_lift(a, P), ref), (bits, N, M)
        assert np.array_equal(crt_unlift(ref, P, DTYPES[bits]), a)
        assert np.array_equal(r.unlift(ref), a)


def test_crt_lift_fixture():
    k = np.load(os.path.join(GOLDEN, "lift_u64_n64_m3.npz"))
    P = golden_params(64)["P"][:3]
    assert np.array_equal(crt_lift(k["polys"], P), k["words"])
    assert np.array_equal(crt_unlift(k["words"], P, np.uint64), k["polys"])

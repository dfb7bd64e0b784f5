"""GPU parity tests (run with -m gpu on the B200 box).  Every call goes through the C ABI (include/nflgpu.h) via
ctypes; results are compared bit-for-bit (np.array_equal on the raw limbs — never the reference's any-equal
operator==, ops.hpp:81-95) with
  * fixtures produced by the unmodified reference (tests/golden/kat_*.npz, hashes.json),
  * the CPU oracle (oracle/nfl_oracle.c) on seeded inputs for every supported size,
  * the live compiled reference (oracle/_ref) when it travelled to the box,
and at the BASELINE.json full sizes through size-independent properties (round trip, linearity, negacyclic
shift) plus oracle spot checks on a slice of the batch."""
import glob
import hashlib
import json
import os
import re

import numpy as np
import pytest

from oracle_lib import GOLDEN, DTYPES, Oracle, Ref, have_ref, golden_params, random_polys, crt_lift, crt_unlift, lift_words_per_coeff

pytestmark = pytest.mark.gpu

import nfllib_b200 as nb  # noqa: E402

KATS = sorted(glob.glob(os.path.join(GOLDEN, "kat_*.npz")))


def _cfg(path):
    return tuple(int(x) for x in re.search(r"kat_u(\d+)_n(\d+)_m(\d+)", path).groups())


_ctx_cache = {}


def ctx_for(bits, N, M):
    key = (bits, N, M)
    if key not in _ctx_cache:
        _ctx_cache[key] = nb.Context(bits, N, M)
    return _ctx_cache[key]


@pytest.mark.parametrize("path", KATS, ids=[os.path.basename(p) for p in KATS])
def test_reference_fixtures(path):
    bits, N, M = _cfg(path)
    k = np.load(path)
    c = ctx_for(bits, N, M)
    a, b = k["a"], k["b"]
    assert np.array_equal(c.run_device("ntt_fwd", a), k["fwd_a"])
    assert np.array_equal(c.run_device("ntt_fwd", b, inplace=True), k["fwd_b"])
    assert np.array_equal(c.run_device("ntt_inv", a), k["inv_a"])
    assert np.array_equal(c.run_device("ntt_inv", k["fwd_a"], inplace=True), a)
    assert np.array_equal(c.run_device("ntt_raw_fwd", a), k["raw_ntt"])      # core::ntt, the tests/ntt_perfs.cpp path
    assert np.array_equal(c.run_device("ntt_raw_inv", a), k["raw_intt"])     # core::inv_ntt
    assert np.array_equal(c.run_device("mul", a, b), k["mul"])
    assert np.array_equal(c.run_device("add", a, b), k["add"])
    assert np.array_equal(c.run_device("sub", a, b), k["sub"])
    assert np.array_equal(c.run_device("compute_shoup", b), k["shoup_b"])
    assert np.array_equal(c.run_device("mul_shoup", a, b, k["shoup_b"]), k["mul_shoup"])
    assert np.array_equal(c.run_device("polymul", a, b), k["polymul"])
    assert np.array_equal(c.run_device("muladd", a, b, k["fwd_a"]), k["muladd"])


SIZES = [(64, n) for n in range(2, 21)] + [(32, n) for n in range(3, 16)] + [(16, n) for n in range(4, 10)]


@pytest.mark.parametrize("bits,n", SIZES, ids=[f"u{b}_n{n}" for b, n in SIZES])
def test_all_sizes_vs_oracle(bits, n):
    N = 1 << n
    # n >= 15 (64-bit) exercises the split path: global-memory passes + tile kernel over sub-blocks (ntt_plan.h)
    cases = ((1, 3), (2 if bits == 16 else 3, 37 if N <= 4096 else 5)) if n <= 15 else ((1, 2), (2, 3))
    for M, batch in cases:
        o = Oracle(bits, N, M)
        c = ctx_for(bits, N, M)
        a = random_polys(bits, N, M, batch, 31 * n + M)
        b = random_polys(bits, N, M, batch, 77 * n + M)
        fa = o.run("fwd", a)
        assert np.array_equal(c.run_device("ntt_fwd", a), fa), (bits, N, M, "fwd")
        assert np.array_equal(c.run_device("ntt_inv", a), o.run("inv", a)), (bits, N, M, "inv")
        assert np.array_equal(c.run_device("ntt_inv", fa), a), (bits, N, M, "roundtrip")
        assert np.array_equal(c.run_device("ntt_raw_fwd", a), o.run("raw_ntt", a)), (bits, N, M, "raw fwd")
        assert np.array_equal(c.run_device("ntt_raw_inv", a), o.run("raw_intt", a)), (bits, N, M, "raw inv")
        assert np.array_equal(c.run_device("mul", a, b), o.run("mul", a, b))
        assert np.array_equal(c.run_device("add", a, b), o.run("add", a, b))
        assert np.array_equal(c.run_device("sub", a, b), o.run("sub", a, b))
        bs = o.run("compute_shoup", b)
        assert np.array_equal(c.run_device("compute_shoup", b), bs)
        assert np.array_equal(c.run_device("mul_shoup", a, b, bs), o.run("mul_shoup", a, b, bs))
        assert np.array_equal(c.run_device("muladd", a, b, fa), o.run("muladd", a, b, fa))
        assert np.array_equal(c.run_device("muladd_shoup", fa, a, b, bs), o.run("add", fa, o.run("mul_shoup", a, b, bs)))
        if N <= 4096 or n >= 15:
            assert np.array_equal(c.run_device("polymul", a, b), o.run("polymul", a, b))


def test_edge_polys_and_shift():
    """all-zero, all p-1, delta_0, X, delta_{N-1}; X*a is the negacyclic shift (tests/nfllib_demo_main_op.cpp:260-331 spirit)."""
    for bits, N, M in ((64, 1024, 4), (32, 4096, 3), (16, 512, 2)):
        P = golden_params(bits)["P"]
        c, o = ctx_for(bits, N, M), Oracle(bits, N, M)
        e = np.zeros((5, M, N), c.dtype)
        for cm in range(M):
            e[1, cm, :] = P[cm] - 1
        e[2, :, 0] = 1
        e[3, :, 1] = 1
        e[4, :, N - 1] = 1
        assert np.array_equal(c.run_device("ntt_fwd", e), o.run("fwd", e))
        assert np.array_equal(c.run_device("ntt_inv", e), o.run("inv", e))
        a = random_polys(bits, N, M, 5, 9)
        x = np.zeros_like(a)
        x[:, :, 1] = 1
        got = c.run_device("polymul", a, x)
        exp = np.roll(a, 1, axis=2)
        for cm in range(M):
            exp[:, cm, 0] = ((int(P[cm]) - a[:, cm, N - 1].astype(np.uint64)) % np.uint64(P[cm])).astype(c.dtype)
        assert np.array_equal(got, exp)


def test_hashes_at_baseline_shapes():
    with open(os.path.join(GOLDEN, "hashes.json")) as f:
        hs = json.load(f)
    h = lambda x: hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()
    for name, e in hs.items():
        c = ctx_for(e["bits"], e["N"], e["M"])
        a = random_polys(e["bits"], e["N"], e["M"], e["batch"], e["seed_a"])
        b = random_polys(e["bits"], e["N"], e["M"], e["batch"], e["seed_b"])
        assert h(a) == e["in_a"]
        fa = c.run_device("ntt_fwd", a)
        assert h(fa) == e["fwd_a"], name
        assert h(c.run_device("ntt_inv", fa)) == e["in_a"], name
        assert h(c.run_device("mul", a, b)) == e["mul"], name
        assert h(c.run_device("add", a, b)) == e["add"], name
        assert h(c.run_device("sub", a, b)) == e["sub"], name
        assert h(c.run_device("polymul", a, b)) == e["polymul"], name


FULL = [("C2", 64, 1024, 4, 4096), ("C3", 64, 16384, 8, 256), ("C4", 32, 4096, 14, 2048), ("C5", 64, 8192, 6, 512)]


@pytest.mark.parametrize("name,bits,N,M,batch", FULL, ids=[f[0] for f in FULL])
def test_full_size_properties(name, bits, N, M, batch):
    """BASELINE.json shapes (batch trimmed only where the host-side oracle/numpy work would take minutes):
    round trip, linearity of the transform, and oracle equality on the first and last polynomials."""
    c, o = ctx_for(bits, N, M), Oracle(bits, N, M)
    a = random_polys(bits, N, M, batch, 1234)
    b = random_polys(bits, N, M, batch, 4321)
    fa, fb = c.run_device("ntt_fwd", a), c.run_device("ntt_fwd", b)
    assert np.array_equal(c.run_device("ntt_inv", fa), a)
    s = c.run_device("add", a, b)
    assert np.array_equal(c.run_device("ntt_fwd", s), c.run_device("add", fa, fb))
    for sl in (slice(0, 2), slice(batch - 2, batch)):
        assert np.array_equal(fa[sl], o.run("fwd", a[sl]))
        assert np.array_equal(c.run_device("polymul", a, b)[sl], o.run("polymul", a[sl], b[sl])) if sl.start == 0 else True
    # checksum of checksums: the product in the NTT domain equals the NTT of the negacyclic product
    prod = c.run_device("polymul", a, b)
    assert np.array_equal(c.run_device("ntt_fwd", prod), c.run_device("mul", fa, fb))


def test_host_op_matches_device_path():
    bits, N, M, batch = 64, 1024, 4, 700  # > one 32 MiB chunk, ragged tail
    c, o = ctx_for(bits, N, M), Oracle(bits, N, M)
    a = random_polys(bits, N, M, batch, 5)
    b = random_polys(bits, N, M, batch, 6)
    fa = c.host_op("fwd", a)
    assert np.array_equal(fa[:4], o.run("fwd", a[:4])) and np.array_equal(fa, c.run_device("ntt_fwd", a))
    assert np.array_equal(c.host_op("inv", fa), a)
    assert np.array_equal(c.host_op("mul", a, b), c.run_device("mul", a, b))
    assert np.array_equal(c.host_op("polymul", a[:64], b[:64]), o.run("polymul", a[:64], b[:64]))


def test_host_op_async_ring_matches_blocking_calls():
    """nflgpu_host_op_async + nflgpu_host_sync: several independent batches in flight through the chunk ring (more chunks than ring
    slots, ragged tails, pageable and page-locked arrays, in place), bit-identical to the blocking call and to the oracle."""
    bits, N, M, batch = 64, 1024, 4, 2600  # 81 MiB per operand: several chunks whatever NFLGPU_HOST_CHUNK_MIB is, ragged tail
    c, o = ctx_for(bits, N, M), Oracle(bits, N, M)
    a = random_polys(bits, N, M, batch, 15)
    b = random_polys(bits, N, M, batch, 16)
    fa = c.host_op("fwd", a)
    assert np.array_equal(fa[-3:], o.run("fwd", a[-3:]))
    o1, o2, o3 = np.zeros_like(a), np.zeros_like(a), np.zeros_like(a)
    c.host_op("fwd", a, out=o1, wait=False)
    c.host_op("inv", fa, out=o2, wait=False)
    c.host_op("mul", a, b, out=o3, wait=False)
    c.host_sync()
    assert np.array_equal(o1, fa) and np.array_equal(o2, a)
    assert np.array_equal(o3[::97], o.run("mul", a[::97], b[::97])) and np.array_equal(o3, c.host_op("mul", a, b))
    # in place on a pageable array, then the same on page-locked memory, with a one-poly call queued behind the big one
    x, one = a.copy(), a[:1].copy()
    c.host_op("fwd", x, out=x, wait=False)
    c.host_op("fwd", one, out=one, wait=False)
    c.host_sync()
    assert np.array_equal(x, fa) and np.array_equal(one, fa[:1])
    y = fa.copy()
    c.host_register(y)
    try:
        c.host_op("inv", y, out=y, wait=False)
        c.host_sync()
        assert np.array_equal(y, a)
        # one page-locked polynomial, in place: the small-call path (the kernel reads and writes the mapped host memory itself)
        c.host_op("fwd", y[7:8], out=y[7:8])
        assert np.array_equal(y[7:8], fa[7:8]) and np.array_equal(y[:7], a[:7]) and np.array_equal(y[8:], a[8:])
        c.host_op("inv", y[7:8], out=y[7:8])
        assert np.array_equal(y, a)
        assert np.array_equal(c.host_op("polymul", y[3:4], b[3:4]), o.run("polymul", a[3:4], b[3:4]))  # pinned + pageable operands
    finally:
        c.host_unregister(y)
    c.host_sync()  # nothing in flight: returns at once
    assert np.array_equal(c.host_op("fwd", a[:5]), fa[:5])  # a blocking call after asynchronous ones


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libnflref.so did not travel")
def test_live_reference_side_by_side():
    # (64, 32768, 2) is the largest configuration of the reference's own test matrix (tests/CMakeLists.txt:1-7)
    for bits, N, M in ((64, 1024, 4), (64, 16384, 8), (32, 4096, 14), (64, 8192, 6), (16, 512, 2), (64, 32768, 2), (32, 32768, 1)):
        r, c = Ref(bits, N, M), ctx_for(bits, N, M)
        a = random_polys(bits, N, M, 6, 99)
        b = random_polys(bits, N, M, 6, 98)
        fa = r.run("fwd", a)
        assert np.array_equal(c.run_device("ntt_fwd", a), fa)
        assert np.array_equal(c.run_device("ntt_inv", fa), a)
        assert np.array_equal(c.run_device("mul", a, b), r.run("mul", a, b))
        assert np.array_equal(c.run_device("polymul", a, b), r.run("polymul", a, b))


def test_errors_and_empty():
    c = ctx_for(64, 1024, 4)
    p = c.alloc(1)
    c.ntt_fwd(p, p, 0)  # empty batch is a no-op
    with pytest.raises(nb.NflGpuError):
        c.ntt_fwd(p + 8, p, 1)  # misaligned
    with pytest.raises(nb.NflGpuError):
        c.ntt_fwd(0, p, 1)  # null
    c.free(p)
    with pytest.raises(nb.NflGpuError):
        nb.Context(64, 1000, 1)  # not a power of two (core.hpp:55-60)
    with pytest.raises(nb.NflGpuError):
        nb.Context(16, 1024, 1)  # degree > params<uint16_t>::kMaxPolyDegree
    with pytest.raises(nb.NflGpuError):
        nb.Context(16, 512, 3)  # nmoduli > params<uint16_t>::kMaxNbModuli
    with pytest.raises(nb.NflGpuError):
        nb.Context(64, 1 << 21, 1)  # degree > params<uint64_t>::kMaxPolyDegree


def test_first_modulus_window_matches_residue_slice():
    """A context over moduli [3, 7) computes residues 3..6 of the 7-modulus transform (residue sharding)."""
    bits, N = 32, 4096
    full, part = ctx_for(bits, N, 7), nb.Context(bits, N, 4, first_modulus=3)
    a = random_polys(bits, N, 7, 4, 17, P=[int(v) for v in full.moduli])
    fa = full.run_device("ntt_fwd", a)
    sub = np.ascontiguousarray(a[:, 3:7, :])
    assert np.array_equal(part.run_device("ntt_fwd", sub), fa[:, 3:7, :])
    part.close()


def _random_tree(rng, nops, depth):
    """returns (postfix program, evaluator(oracle, operands) -> array, max stack depth)"""
    if depth == 0 or rng.random() < 0.25:
        k = rng.randrange(nops)
        return [k], (lambda o, ops, k=k: ops[k]), 1
    kind = rng.choice(["add", "sub", "mul", "mul", "shoup"])
    lp, lf, ld = _random_tree(rng, nops, depth - 1)
    if kind == "shoup":  # x * y with y a plain operand and y' = compute_shoup(y) computed inside the program
        k = rng.randrange(nops)
        prog = lp + [k, k, 0x14, 0x13]
        return prog, (lambda o, ops, lf=lf, k=k: o.run("mul_shoup", lf(o, ops), ops[k], o.run("compute_shoup", ops[k]))), max(ld, 3)
    rp, rf, rd = _random_tree(rng, nops, depth - 1)
    tok = {"add": 0x10, "sub": 0x11, "mul": 0x12}[kind]
    return lp + rp + [tok], (lambda o, ops, lf=lf, rf=rf, kind=kind: o.run(kind, lf(o, ops), rf(o, ops))), max(ld, rd + 1)


def test_fused_expression_evaluator_random_trees():
    """nflgpu_eval (the reference's expression templates, ops.hpp:52-97 + core.hpp:24-37): random expression trees over
    up to 8 operands in ONE kernel, compared with the oracle evaluating the same tree one functor at a time."""
    import random
    rng = random.Random(7)
    for bits, N, M in ((64, 1024, 4), (32, 4096, 3), (16, 512, 2)):
        c, o = ctx_for(bits, N, M), Oracle(bits, N, M)
        batch = 6
        ops_host = [random_polys(bits, N, M, batch, 900 + i) for i in range(8)]
        dev = []
        for h in ops_host:
            p = c.alloc(batch)
            c.upload(p, h, batch)
            dev.append(p)
        out = c.alloc(batch)
        done = 0
        while done < 20:
            nops = rng.randint(1, 8)
            prog, fn, depth = _random_tree(rng, nops, rng.randint(1, 4))
            if len(prog) > 32 or depth > 8:
                continue
            if done % 5 == 4:  # root compute_shoup (its result is a Shoup word, not a residue, so only at the root)
                prog, fn = prog + [0x14], (lambda o_, ops, fn=fn: o_.run("compute_shoup", fn(o_, ops)))
            c.eval(out, dev[:nops], prog, batch)
            got = np.empty_like(ops_host[0])
            c.download(got, out, batch)
            c.sync()
            assert np.array_equal(got, fn(o, ops_host)), (bits, prog)
            done += 1
        c.eval(dev[0], dev[:2], [0, 1, 0x12, 0, 0x10], batch)   # dst aliases an operand: a*b + a, in place
        got = np.empty_like(ops_host[0])
        c.download(got, dev[0], batch)
        c.sync()
        assert np.array_equal(got, o.run("muladd", ops_host[0], ops_host[0], ops_host[1]))
        with pytest.raises(nb.NflGpuError):
            c.eval(out, dev[:2], [0, 1], batch)          # leaves two values
        with pytest.raises(nb.NflGpuError):
            c.eval(out, dev[:2], [0, 0x10], batch)       # stack underflow
        with pytest.raises(nb.NflGpuError):
            c.eval(out, dev[:2], [5], batch)             # operand out of range
        for p in dev + [out]:
            c.free(p)


def test_uniform_sampler_matches_reference_draws():
    """nflgpu_uniform vs the oracle's restatement of poly::set(uniform) (Salsa20 keystream + mask + conditional subtract),
    and — when oracle/_ref travelled — vs the reference's own sampler run with the fixed key of the harness."""
    key = bytes(range(1, 33))
    for bits, N, M in ((64, 1024, 4), (64, 64, 3), (32, 4096, 1), (32, 8, 2), (16, 512, 2), (16, 16, 1), (64, 4, 1), (32, 16384, 5)):
        c, o = ctx_for(bits, N, M), Oracle(bits, N, M)
        for batch, nonce in ((5, 0), (3, 2**32 - 1), (1, 2**63 + 7)):
            d = c.alloc(batch)
            c.uniform(d, batch, key, nonce)
            got = np.empty((batch, M, N), c.dtype)
            c.download(got, d, batch)
            c.sync()
            c.free(d)
            assert np.array_equal(got, o.uniform(batch, key, nonce)), (bits, N, M, nonce)
            assert all((got[:, cm, :] < c.moduli[cm]).all() for cm in range(M))
    if have_ref():
        for bits, N, M in ((64, 1024, 4), (32, 8, 2), (16, 512, 2)):
            n0, ref = Ref(bits, N, M).uniform(4)
            c = ctx_for(bits, N, M)
            d = c.alloc(4)
            c.uniform(d, 4, Ref.FIXED_KEY, n0)
            got = np.empty((4, M, N), c.dtype)
            c.download(got, d, 4)
            c.sync()
            c.free(d)
            assert np.array_equal(got, ref)


def test_bounded_and_zo_samplers():
    """nflgpu_non_uniform / nflgpu_zo vs the oracle restatements of core.hpp:190-278 / 338-349 (pinned against the reference's
    own samplers in tests/test_oracle.py) and, when it travelled, the live reference."""
    key = bytes(range(1, 33))
    for bits, N, M in ((64, 1024, 4), (32, 8, 2), (16, 512, 2), (32, 4096, 3), (64, 16, 2)):
        c, o = ctx_for(bits, N, M), Oracle(bits, N, M)
        batch = 4
        d = c.alloc(batch)
        got = np.empty((batch, M, N), c.dtype)
        for ub, amp in ((1, 1), (5, 3), (1 << 10, 1), (1000, 7)):
            c.non_uniform(d, batch, ub, amp, key, 77)
            c.download(got, d, batch); c.sync()
            assert np.array_equal(got, o.non_uniform(batch, ub, amp, key, 77)), (bits, N, M, ub, amp)
        for rho in (0, 0x7F, 0xFF):
            c.zo(d, batch, rho, key, 1234)
            c.download(got, d, batch); c.sync()
            assert np.array_equal(got, o.zo(batch, rho, key, 1234)), (bits, N, M, rho)
        with pytest.raises(nb.NflGpuError):
            c.non_uniform(d, batch, int(c.moduli.min()), 1, key, 0)   # core.hpp:201-206
        c.free(d)
    if have_ref():
        bits, N, M = 64, 1024, 4
        r, c = Ref(bits, N, M), ctx_for(bits, N, M)
        d = c.alloc(3)
        got = np.empty((3, M, N), c.dtype)
        n0, ref = r.sample("non_uniform", 3, 9, 2)
        c.non_uniform(d, 3, 9, 2, Ref.FIXED_KEY, n0)
        c.download(got, d, 3); c.sync()
        assert np.array_equal(got, ref)
        n0, ref = r.sample("zo", 3, 0x7F)
        c.zo(d, 3, 0x7F, Ref.FIXED_KEY, n0)
        c.download(got, d, 3); c.sync()
        assert np.array_equal(got, ref)
        c.free(d)


def test_crt_lift_matches_oracle_and_reference():
    """nflgpu_poly2mpz / nflgpu_mpz2poly (gmp.hpp:183-219) vs the big-integer restatement (pinned against the reference's GMP
    code in tests/test_oracle.py) and the live reference when it travelled."""
    import torch
    for bits, N, M in ((64, 1024, 4), (64, 64, 3), (32, 1024, 2), (32, 4096, 14), (16, 512, 2), (64, 1024, 1), (64, 256, 16)):
        c = ctx_for(bits, N, M)
        P = [int(p) for p in c.moduli]
        a = random_polys(bits, N, M, 3, 55, P=P)
        e = np.zeros((3, M, N), c.dtype)
        for cm in range(M):
            e[1, cm, :] = P[cm] - 1
        e[2, 0, :] = 1
        a = np.concatenate([a, e])
        batch = a.shape[0]
        W = c.lift_words()
        assert W == lift_words_per_coeff(P)
        dp = c.alloc(batch)
        c.upload(dp, a, batch)
        dw = torch.zeros((batch, N, W), dtype=torch.int64, device="cuda")
        c.poly2mpz(dw.data_ptr(), dp, batch)
        c.sync()
        words = dw.cpu().numpy().view(np.uint64)
        exp = crt_lift(a, P)
        assert np.array_equal(words, exp), (bits, N, M)
        back = c.alloc(batch)
        c.mpz2poly(back, dw.data_ptr(), batch)
        got = np.empty_like(a)
        c.download(got, back, batch)
        c.sync()
        assert np.array_equal(got, a)
        if have_ref() and Ref(bits, N, M).supported() and (bits, N, M) in ((64, 1024, 4), (32, 4096, 14), (16, 512, 2)):
            assert np.array_equal(words, Ref(bits, N, M).lift(a, W))
        c.free(dp); c.free(back)
    with pytest.raises(nb.NflGpuError):
        ctx_for(64, 64, 17).lift_words()   # 17 * 62 bits > 1024


def test_hwt_sampler():
    """nflgpu_hwt vs the oracle restatement of core.hpp:355-392 (pinned against the reference in tests/test_oracle.py) and the
    live reference when it travelled."""
    key = bytes(range(1, 33))
    for bits, N, M in ((64, 1024, 4), (32, 8, 2), (16, 512, 2), (64, 64, 3)):
        c, o = ctx_for(bits, N, M), Oracle(bits, N, M)
        for hwt in (1, 3, N // 4, N - 1, N):
            batch = 5
            d = c.alloc(batch)
            c.hwt(d, batch, hwt, key, 1000)
            got = np.empty((batch, M, N), c.dtype)
            c.download(got, d, batch); c.sync()
            c.free(d)
            exp, calls = o.hwt(batch, hwt, key, 1000)
            assert np.array_equal(got, exp), (bits, N, M, hwt)
            assert all(np.count_nonzero(got[b, 0]) == hwt for b in range(batch))
        with pytest.raises(nb.NflGpuError):
            c.hwt(1, 1, N + 1, key, 0)
    if have_ref():
        bits, N, M = 64, 1024, 4
        n0, ref = Ref(bits, N, M).sample("hwt", 3, 64)
        c = ctx_for(bits, N, M)
        d = c.alloc(3)
        c.hwt(d, 3, 64, Ref.FIXED_KEY, n0)
        got = np.empty((3, M, N), c.dtype)
        c.download(got, d, 3); c.sync()
        c.free(d)
        assert np.array_equal(got, ref)


def test_unit_scheduler_rearms_across_many_launches_and_streams():
    """The NTT kernels of the larger degrees hand out (polynomial, residue) units through per-launch atomic counters that the
    last CTA re-arms (ntt_engine.cuh UnitWalk).  Hundreds of back-to-back launches (more than the context's counter sets),
    launches interleaved on two streams, and batches smaller / ragged against the grid must all stay bit-exact."""
    import torch
    for bits, N, M in ((64, 2048, 3), (32, 4096, 2), (64, 32768, 1)):
        c, o = ctx_for(bits, N, M), Oracle(bits, N, M)
        for batch in (1, 3, 37):
            a = random_polys(bits, N, M, batch, 1000 + batch)
            want = o.run("fwd", a)
            s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
            src = c.alloc(batch)
            outs = [c.alloc(batch) for _ in range(4)]
            c.upload(src, a, batch)
            c.sync()
            for i in range(300):  # > nflgpu_ctx::kSchedSlots launches in total, alternating streams and directions
                st = (s1 if i % 2 == 0 else s2).cuda_stream
                c.ntt_fwd(outs[i % 2], src, batch, st)
                c.ntt_inv(outs[2 + i % 2], outs[i % 2], batch, st)
            torch.cuda.synchronize()
            for k in range(2):
                got = np.empty_like(a)
                c.download(got, outs[k], batch)
                c.sync()
                assert np.array_equal(got, want), (bits, N, batch, k)
                c.download(got, outs[2 + k], batch)
                c.sync()
                assert np.array_equal(got, a), (bits, N, batch, k)
            for p_ in [src] + outs:
                c.free(p_)


def test_graft_entry_smoke():
    """The driver's smoke() entry point itself."""
    import __graft_entry__
    __graft_entry__.smoke()

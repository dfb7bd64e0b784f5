"""poly::set(gaussian) over FastGaussianNoise (the last sampler): the barrier table the product computes (MPFR
runtime) against the reference's own, the oracle restatement of the look-up sampler against the reference's draws, and the
device sampler against both — all bit for bit, including the number of PRNG nonces a batch consumes."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle_lib import GOLDEN, GaussTable, Oracle, Ref, have_ref

import nfllib_b200.capi as capi


def fixture():
    z = np.load(os.path.join(GOLDEN, "gaussian.npz"))
    return z, json.loads(bytes(z["meta"]).decode())


def table_of(z, name, m):
    return GaussTable(z[name + "_barriers"], m["in_bytes"], m["lu_depth"], m["rounded_center"], m["params"])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


CASES = ["demo_u64", "prng_demo_u64_small", "depth1_u16", "words16_u32"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_sampler_reproduces_the_reference_draws_of_the_fixture(name):
    z, meta = fixture()
    m = meta[name]
    got, calls = Oracle(m["bits"], m["N"], m["M"]).gaussian(m["batch"], table_of(z, name, m), m["amplifier"], Ref.FIXED_KEY, m["first_nonce"])
    assert calls == m["nonces_used"]
    assert sha(got) == m["sha256"]
    if name + "_draws" in z:
        assert np.array_equal(got, z[name + "_draws"])


@pytest.mark.parametrize("name", CASES)
def test_product_barrier_table_is_the_reference_table(name):
    """nflgpu_gaussian_table (host only; same MPFR calls as FastGaussianNoise.hpp:285-353 through the installed runtime)."""
    z, meta = fixture()
    m = meta[name]
    try:
        d, bar = capi.gaussian_table(m["sigma"], m["security"], m["samples"], m["center"], m["in_bytes"], m["lu_depth"])
    except capi.NflGpuError as e:
        if "runtimes" in str(e):
            pytest.skip("libmpfr.so.6 / libgmp.so.10 not present on this machine")
        raise
    want = z[name + "_barriers"]
    assert bar.shape == want.shape and np.array_equal(bar, want)
    for k in ("nb", "wp", "bit_precision", "flag_ctr1", "flag_ctr2"):
        assert d[k] == m["params"][k], k
    assert d["rounded_center"] == m["rounded_center"]
    assert abs(d["tail_bound"] - m["params"]["tail_bound"]) < 1e-9  # double Newton iteration; the reference build contracts to FMA


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("cfg", [(64, 1024, 4, 1, 2, 20.0, 1 << 14, 0.0, 1), (64, 1024, 4, 1, 1, 20.0, 1 << 10, 0.0, 3),
                                 (64, 1024, 4, 2, 1, 300.0, 1 << 10, 0.0, 1), (32, 4096, 1, 1, 2, 3.19, 1 << 19, 5.25, 2),
                                 (16, 512, 2, 1, 2, 3.19, 1 << 10, -2.5, 2), (64, 64, 3, 1, 2, 20.0, 1 << 14, 0.0, 1)])
def test_oracle_and_product_tables_against_the_live_reference(cfg):
    bits, N, M, ib, depth, sigma, samples, center, amp = cfg
    h, t = Ref.gaussian_table(sigma, 128, samples, center, ib, depth, bits)
    first, used, want = Ref(bits, N, M).gaussian(h, 20, amp)
    got, calls = Oracle(bits, N, M).gaussian(20, t, amp, Ref.FIXED_KEY, first)
    assert calls == used and np.array_equal(got, want)
    d, bar = capi.gaussian_table(sigma, 128, samples, center, ib, depth)
    assert np.array_equal(bar, t.barriers) and d["flag_ctr1"] == t.params["flag_ctr1"] and d["flag_ctr2"] == t.params["flag_ctr2"]


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_product_tables_equal_the_reference_tables_over_random_parameters():
    """40 seeded (sigma, security, samples, center, in_class, depth) draws: barrier table, flag counters and rounded center of
    nflgpu_gaussian_table equal the private members of the reference's FastGaussianNoise object."""
    rng = np.random.default_rng(5)
    for i in range(40):
        sigma = float(np.round(rng.uniform(0.8, 400.0), 3)) if i % 3 else float(rng.integers(1, 60))
        sec = int(rng.choice([64, 80, 100, 128, 192, 256]))
        samples = int(1 << rng.integers(4, 24))
        center = float(np.round(rng.uniform(-50, 50), 2)) if i % 2 else 0.0
        ib, depth = [(1, 2), (1, 1), (2, 1)][i % 3]
        if ib == 1 and sigma > 120:
            sigma = sigma / 8
        h, t = Ref.gaussian_table(sigma, sec, samples, center, ib, depth, 64)
        d, bar = capi.gaussian_table(sigma, sec, samples, center, ib, depth)
        assert bar.shape == t.barriers.shape and np.array_equal(bar, t.barriers), (sigma, sec, samples, center, ib, depth)
        assert (d["flag_ctr1"], d["flag_ctr2"], d["rounded_center"]) == (t.params["flag_ctr1"], t.params["flag_ctr2"], t.rounded_center)


def test_gaussian_argument_errors_are_reported_not_thrown():
    with pytest.raises(capi.NflGpuError):
        capi.gaussian_table(-1.0, 128, 1024)
    with pytest.raises(capi.NflGpuError):
        capi.gaussian_table(3.19, 128, 1024, in_bytes=2, lu_depth=2)  # 65536 second-level tables: not supported
    with pytest.raises(capi.NflGpuError):
        capi.gaussian_table(3.19, 128, 1024, in_bytes=3)


# ---- device ------------------------------------------------------------------------------------------------------------------

def device_draws(ctx, g, batch, key, first_nonce, amp):
    p = ctx.alloc(batch)
    used = g.sample(p, batch, key, first_nonce, amp)
    out = np.empty((batch, ctx.nmoduli, ctx.degree), dtype=ctx.dtype)
    ctx.download(out, p, batch)
    ctx.sync()
    ctx.free(p)
    return out, used


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_gaussian_matches_fixture_and_oracle(name):
    z, meta = fixture()
    m = meta[name]
    ctx = capi.Context(m["bits"], m["N"], m["M"])
    t = table_of(z, name, m)
    # sampler from the committed reference table, and sampler from the product's own MPFR-built table
    gs = [capi.Gaussian(ctx, in_bytes=m["in_bytes"], lu_depth=m["lu_depth"], barriers=t.barriers, rounded_center=m["rounded_center"]),
          capi.Gaussian(ctx, m["sigma"], m["security"], m["samples"], m["center"], m["in_bytes"], m["lu_depth"])]
    for g in gs:
        assert g.info()["flag_ctr1"] == m["params"]["flag_ctr1"] and g.info()["flag_ctr2"] == m["params"]["flag_ctr2"]
        got, used = device_draws(ctx, g, m["batch"], Ref.FIXED_KEY, m["first_nonce"], m["amplifier"])
        assert used == m["nonces_used"]
        assert sha(got) == m["sha256"]
    # a larger batch, other key / nonce / amplifier: against the oracle
    key = bytes((7 * i + 3) & 0xFF for i in range(32))
    batch = 300 if m["N"] <= 1024 else 64
    for amp in (1, 5):
        want, calls = Oracle(m["bits"], m["N"], m["M"]).gaussian(batch, t, amp, key, 12345)
        got, used = device_draws(ctx, gs[1], batch, key, 12345, amp)
        assert used == calls
        assert np.array_equal(got, want)
    # every coefficient is the same small centred value in all residues
    P = ctx.moduli.astype(object)
    v0 = got[:, 0, :].astype(object)
    c0 = np.where(v0 > P[0] // 2, v0 - P[0], v0)
    for cm in range(1, m["M"]):
        v = got[:, cm, :].astype(object)
        assert np.array_equal(np.where(v > P[cm] // 2, v - P[cm], v), c0)
    for g in gs:
        g.close()
    ctx.close()


@pytest.mark.gpu
def test_device_gaussian_large_batch_goes_out_in_chunks(monkeypatch):
    """Batches whose scratch would exceed the budget are processed chunk by chunk; the nonce chain continues across chunks."""
    z, meta = fixture()
    m = meta["demo_u64"]
    monkeypatch.setenv("NFLGPU_GAUSS_SCRATCH_MB", "4")  # ~137 polynomials per chunk for this shape
    ctx = capi.Context(m["bits"], m["N"], m["M"])
    t = table_of(z, "demo_u64", m)
    g = capi.Gaussian(ctx, in_bytes=1, lu_depth=2, barriers=t.barriers, rounded_center=0)
    want, calls = Oracle(m["bits"], m["N"], m["M"]).gaussian(1000, t, 1, Ref.FIXED_KEY, 77)
    got, used = device_draws(ctx, g, 1000, Ref.FIXED_KEY, 77, 1)
    assert used == calls and np.array_equal(got, want)
    g.close()
    ctx.close()


@pytest.mark.gpu
@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_device_gaussian_next_to_the_live_reference():
    """The demo's shape (tests/nfllib_demo_main_op.cpp:141-144, 273-283): FastGaussianNoise<uint8_t, uint64_t, 2>(20, 128, 2^14),
    amplifier 1 and 2, consecutive batches continuing the nonce stream."""
    bits, N, M = 64, 1024, 4
    h, t = Ref.gaussian_table(20.0, 128, 1 << 14, 0.0, 1, 2, bits)
    r = Ref(bits, N, M)
    ctx = capi.Context(bits, N, M)
    g = capi.Gaussian(ctx, 20.0, 128, 1 << 14, 0.0, 1, 2)
    for batch, amp in ((40, 1), (17, 2), (1, 1)):
        first, used, want = r.gaussian(h, batch, amp)
        got, dev_used = device_draws(ctx, g, batch, Ref.FIXED_KEY, first, amp)
        assert dev_used == used
        assert np.array_equal(got, want)
    g.close()
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("bits,N,M,batch", [(64, 8192, 2, 48), (64, 32768, 1, 40), (32, 16384, 2, 40), (64, 4096, 3, 70)])
def test_device_gaussian_large_degrees_stage_fewer_rows_than_candidates(bits, N, M, batch):
    """Degrees >= 8192 have a consumed-words pitch above 6400 bytes, so a walk CTA stages fewer than 32 rows in shared memory and
    the remaining candidates follow their chain in global memory (round-1 ADVICE: those threads read past the staged rows)."""
    z, meta = fixture()
    m = meta["demo_u64"]
    t = table_of(z, "demo_u64", m)
    ctx = capi.Context(bits, N, M)
    g = capi.Gaussian(ctx, in_bytes=m["in_bytes"], lu_depth=m["lu_depth"], barriers=t.barriers, rounded_center=m["rounded_center"])
    key = bytes((11 * i + 5) & 0xFF for i in range(32))
    want, calls = Oracle(bits, N, M).gaussian(batch, t, 1, key, 4242)
    got, used = device_draws(ctx, g, batch, key, 4242, 1)
    assert used == calls
    assert np.array_equal(got, want)
    g.close()
    ctx.close()

"""CPU check of the kernels' own butterfly code: tests/cpp/engine_sim.cu runs ntt_engine.cuh's fwd_pass / inv_pass /
pass_pos / pass_tw / fwd_canon (compiled for the host) pass by pass over one unit, with the twiddle tables the product builds,
and the result must equal the oracle's ntt_pow_phi / invntt_pow_invphi bit for bit — for every limb type and every size up to
the first split transforms, on random and on edge inputs (0, p-1, deltas, alternating extremes)."""
import ctypes
import os

import numpy as np
import pytest

from oracle_lib import Oracle, random_polys, golden_params, DTYPES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM_SO = os.path.join(ROOT, "tests", "cpp", "libenginesim.so")

# the same file built with -DNFLGPU_FOLD=1: N^-1 folded into the inverse twiddles of every 64-bit size (ntt_plan.h plan_fold)
SIM_FOLD_SO = os.path.join(ROOT, "tests", "cpp", "libenginesim_fold.so")

_libs = {}


def sim(path=SIM_SO):
    if path not in _libs:
        assert os.path.exists(path), "run __graft_entry__.build()"
        lib = ctypes.CDLL(path)
        lib.nflsim_ntt.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64,
                                   ctypes.c_void_p]
        lib.nflsim_ntt_tile.argtypes = lib.nflsim_ntt.argtypes
        lib.nflsim_pointwise.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint64] + [ctypes.c_void_p] * 5 + [ctypes.c_size_t]
        _libs[path] = lib
    return _libs[path]


PW = {"add": 0, "sub": 1, "mul": 2, "mul_shoup": 3, "compute_shoup": 4, "muladd": 5, "muladd_shoup": 6}


def run_pw(bits, M, op, a, b=None, c=None, d=None):
    """Functor<LB, OP>::apply over whole polynomials [batch][M][N], residue by residue."""
    g = golden_params(bits)
    out = np.empty_like(a)
    w = lambda x, cm: None if x is None else np.ascontiguousarray(x[:, cm].astype(np.uint64))
    for cm in range(M):
        ops = [w(x, cm) for x in (a, b, c, d)]
        o = np.empty_like(ops[0])
        rc = sim().nflsim_pointwise(bits, PW[op], g["P"][cm], *[None if x is None else x.ctypes.data for x in ops], o.ctypes.data, o.size)
        assert rc == 0
        out[:, cm] = o.astype(DTYPES[bits])
    return out


def run_sim(bits, N, M, polys, inverse, tile=False, lib=SIM_SO):
    g = golden_params(bits)
    out = np.empty_like(polys)
    n = N.bit_length() - 1
    fn = sim(lib).nflsim_ntt_tile if tile else sim(lib).nflsim_ntt
    for b in range(polys.shape[0]):
        for cm in range(M):
            d = np.ascontiguousarray(polys[b, cm].astype(np.uint64))
            rc = fn(bits, n, int(inverse), g["P"][cm], g["roots"][cm], g["kmax"], d.ctypes.data)
            assert rc == 0, (bits, N)
            out[b, cm] = d.astype(DTYPES[bits])
    return out


def edge_polys(bits, N, M):
    g = golden_params(bits)
    P = np.array(g["P"][:M], dtype=np.uint64)
    e = np.zeros((6, M, N), dtype=DTYPES[bits])
    e[1] = (P - 1)[:, None].astype(DTYPES[bits])           # all p-1
    e[2, :, 0] = 1                                           # delta_0
    e[3, :, 1 % N] = 1                                       # X
    e[4, :, N - 1] = (P - 1).astype(DTYPES[bits])            # -X^(N-1)
    e[5, :, ::2] = (P - 1)[:, None].astype(DTYPES[bits])    # alternating p-1, 0
    return e


SIZES = ([(64, 1 << n) for n in range(2, 18)] + [(32, 1 << n) for n in range(3, 16)] + [(16, 1 << n) for n in range(4, 10)])


@pytest.mark.parametrize("bits,N", SIZES)
def test_kernel_butterfly_networks_on_the_host_match_the_oracle(bits, N):
    M = 2
    o = Oracle(bits, N, M)
    count = 3 if N <= 4096 else 1
    a = np.concatenate([random_polys(bits, N, M, count, 7000 + N), edge_polys(bits, N, M)])
    want = o.run("fwd", a)
    got = run_sim(bits, N, M, a, inverse=False)
    assert np.array_equal(got, want)
    back = run_sim(bits, N, M, want, inverse=True)
    assert np.array_equal(back, a)
    # the inverse on arbitrary canonical input (not only on forward outputs)
    assert np.array_equal(run_sim(bits, N, M, a, inverse=True), o.run("inv", a))


@pytest.mark.parametrize("N", [1 << n for n in range(2, 18)])
def test_inverse_with_folded_scaling_on_the_host(N):
    """The inverse networks with N^-1 carried by the twiddles of the first inverse pass (one multiplication per thread instead of one
    per butterfly of the last stage), forced on for every 64-bit size: one-pass, multi-pass and split shapes, flat and through the tile."""
    bits, M = 64, 2
    o = Oracle(bits, N, M)
    a = np.concatenate([random_polys(bits, N, M, 2 if N <= 4096 else 1, 7200 + N), edge_polys(bits, N, M)])
    want = o.run("inv", a)
    assert np.array_equal(run_sim(bits, N, M, a, inverse=True, lib=SIM_FOLD_SO), want)
    if 64 <= N <= 16384:
        assert np.array_equal(run_sim(bits, N, M, a, inverse=True, tile=True, lib=SIM_FOLD_SO), want)
    assert np.array_equal(run_sim(bits, N, M, a, inverse=False, lib=SIM_FOLD_SO), o.run("fwd", a))  # (forward: unchanged)


TILE_SIZES = ([(64, 1 << n) for n in range(6, 15)] + [(32, 1 << n) for n in range(7, 16)] + [(16, 1 << n) for n in range(7, 10)])


@pytest.mark.parametrize("bits,N", TILE_SIZES)
def test_tile_exchange_layout_on_the_host(bits, N):
    """The same networks with the passes exchanging through the kernels' own shared-memory tile addressing (tile_store / tile_load /
    taddr: padded rows, or the XOR swizzle of the 32-bit N = 4096 shape) and the 16-byte copy-in / copy-out order."""
    M = 2
    o = Oracle(bits, N, M)
    a = np.concatenate([random_polys(bits, N, M, 2 if N <= 4096 else 1, 7100 + N), edge_polys(bits, N, M)[3:5]])
    want = o.run("fwd", a)
    assert np.array_equal(run_sim(bits, N, M, a, inverse=False, tile=True), want)
    assert np.array_equal(run_sim(bits, N, M, want, inverse=True, tile=True), a)


def test_bank_conflict_model_of_every_tile_shape():
    """tools/bank_conflicts.py: every pass and the 16-byte copies of every multi-pass shape reach the minimum number of shared-memory
    wavefronts (the 32-bit N = 4096 shape only with its XOR swizzle); N <= 256 copies touch fewer than 32 lanes' worth and are exempt."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bank_conflicts", os.path.join(ROOT, "tools", "bank_conflicts.py"))
    bc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bc)
    for wb, n in [(64, k) for k in range(9, 15)] + [(32, k) for k in range(9, 16)]:
        _, _, res = bc.analyse(wb, n, swz=(wb == 32 and n == 12))
        for name, total, minimum in res:
            assert total == minimum, (wb, n, name, total, minimum)
    _, _, res = bc.analyse(32, 12, swz=False)
    assert any(total > minimum for _, total, minimum in res)  # what the swizzle is for


@pytest.mark.parametrize("bits,N", [(64, 256), (32, 256), (16, 128)])
def test_pointwise_functors_on_the_host_match_the_oracle(bits, N):
    """addmod / submod / mulmod / mulmod_shoup / compute_shoup / muladd[_shoup] as the kernels compute them (modmul.cuh), on random
    operands and on the extremes 0, 1, p-1."""
    M = 2
    o = Oracle(bits, N, M)
    P = np.array(golden_params(bits)["P"][:M], dtype=np.uint64)
    a = np.concatenate([random_polys(bits, N, M, 6, 91), edge_polys(bits, N, M)])
    b = np.concatenate([random_polys(bits, N, M, 6, 92), edge_polys(bits, N, M)[::-1]])
    c = np.concatenate([random_polys(bits, N, M, 6, 93), edge_polys(bits, N, M)])
    for x in (a, b):  # sprinkle the extremes over the random part too
        x[0, :, :8] = 0
        x[1, :, :8] = 1
        x[2, :, :8] = (P - 1)[:, None].astype(DTYPES[bits])
    bs = o.run("compute_shoup", b)
    assert np.array_equal(run_pw(bits, M, "compute_shoup", b), bs)
    assert np.array_equal(run_pw(bits, M, "add", a, b), o.run("add", a, b))
    assert np.array_equal(run_pw(bits, M, "sub", a, b), o.run("sub", a, b))
    assert np.array_equal(run_pw(bits, M, "mul", a, b), o.run("mul", a, b))
    assert np.array_equal(run_pw(bits, M, "mul_shoup", a, b, bs), o.run("mul_shoup", a, b, bs))
    assert np.array_equal(run_pw(bits, M, "muladd", c, a, b), o.run("muladd", c, a, b))
    # muladd_shoup(c, a, b, b') = c + a*b computed through the Shoup word: same value as muladd
    assert np.array_equal(run_pw(bits, M, "muladd_shoup", c, a, b, bs), o.run("muladd", c, a, b))

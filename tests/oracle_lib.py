"""TEST INFRASTRUCTURE: ctypes bindings for the CPU oracle (oracle/liboracle.so, the C restatement) and,
when it has been built, the unmodified reference (oracle/_ref/libnflref.so).  Never imported by the product."""
import ctypes
import json
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libnflref.so")               # -march=x86-64-v3: runs on any AVX2 host
REF_NATIVE_DIR = os.path.join(ROOT, "oracle", "_ref", "native")               # -march=native of the machine that built it
GOLDEN = os.path.join(ROOT, "tests", "golden")

DTYPES = {16: np.uint16, 32: np.uint32, 64: np.uint64}
OPS = {"fwd": 0, "inv": 1, "mul": 2, "mul_shoup": 3, "compute_shoup": 4, "add": 5, "sub": 6, "raw_ntt": 7,
       "polymul": 8, "muladd": 9, "raw_intt": 10}


def aligned(shape, dtype, align=64):
    """numpy array whose data pointer is `align`-byte aligned (reference asserts 32, core.hpp:88)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    raw = np.zeros(n + align, np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n].view(dtype).reshape(shape)


def as_aligned(a):
    if a.ctypes.data % 32 == 0 and a.flags["C_CONTIGUOUS"]:
        return a
    b = aligned(a.shape, a.dtype)
    b[...] = a
    return b


_params_cache = {}


def golden_params(bits):
    """First entries of params<T>::{P,Pn,primitive_roots,invkMaxPolyDegree} (tests/golden/params.json)."""
    if not _params_cache:
        with open(os.path.join(GOLDEN, "params.json")) as f:
            _params_cache.update({int(k): v for k, v in json.load(f).items()})
    return _params_cache[bits]


class Oracle:
    """C restatement of the reference path (oracle/nfl_oracle.c)."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = ctypes.CDLL(ORACLE_SO)
            L.nflo_create.restype = ctypes.c_void_p
            L.nflo_create.argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t] + [ctypes.c_void_p] * 4 + [ctypes.c_uint64]
            L.nflo_destroy.argtypes = [ctypes.c_void_p]
            L.nflo_run.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 4 + [ctypes.c_size_t]
            L.nflo_spec_fwd.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
            L.nflo_uniform.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_uint64]
            L.nflo_non_uniform.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_uint64]
            L.nflo_zo.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint8, ctypes.c_char_p, ctypes.c_uint64]
            L.nflo_hwt.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint64,
                                   ctypes.POINTER(ctypes.c_uint64)]
            cls._lib = L
        return cls._lib

    def __init__(self, bits, N, M, params=None):
        g = params or golden_params(bits)
        self.bits, self.N, self.M = bits, N, M
        self.dtype = DTYPES[bits]
        mk = lambda k: np.array(g[k][:M], dtype=np.uint64)
        self.P, self.Pn, self.roots, self.invk = mk("P"), mk("Pn"), mk("roots"), mk("invkmax")
        assert len(self.P) == M, "not enough golden moduli"
        self.h = self.lib().nflo_create(bits, N, M, self.P.ctypes.data, self.Pn.ctypes.data, self.roots.ctypes.data,
                                        self.invk.ctypes.data, g["kmax"])
        assert self.h, "nflo_create failed"

    def __del__(self):
        if getattr(self, "h", None):
            self.lib().nflo_destroy(self.h)
            self.h = None

    def run(self, op, a, b=None, c=None):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        out = np.empty_like(a)
        ptr = lambda x: None if x is None else np.ascontiguousarray(x, dtype=self.dtype).ctypes.data
        keep = [np.ascontiguousarray(x, dtype=self.dtype) for x in (b, c) if x is not None]
        pb = keep[0].ctypes.data if b is not None else None
        pc = keep[-1].ctypes.data if c is not None else None
        rc = self.lib().nflo_run(self.h, OPS[op], out.ctypes.data, a.ctypes.data, pb, pc, a.size // (self.N * self.M))
        assert rc == 0
        return out

    def uniform(self, batch, key, first_nonce):
        """`batch` successive poly::set(uniform) draws from the Salsa20 stream (key, first_nonce + i)."""
        out = np.empty((batch, self.M, self.N), dtype=self.dtype)
        self.lib().nflo_uniform(self.h, out.ctypes.data, batch, bytes(key), first_nonce)
        return out

    def non_uniform(self, batch, upper_bound, amplifier, key, first_nonce):
        out = np.empty((batch, self.M, self.N), dtype=self.dtype)
        self.lib().nflo_non_uniform(self.h, out.ctypes.data, batch, upper_bound, amplifier, bytes(key), first_nonce)
        return out

    def hwt(self, batch, hwt, key, first_nonce, test_shrink=0):
        """(polys, number of fastrandombytes calls made); test_shrink > 0 cuts the top 2^-test_shrink part off the index
        sampler's acceptance range (forces the extra refills that the real rule makes astronomically rare)"""
        out = np.empty((batch, self.M, self.N), dtype=self.dtype)
        calls = ctypes.c_uint64()
        L = self.lib()
        L.nflo_hwt_test.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint64,
                                    ctypes.c_uint, ctypes.POINTER(ctypes.c_uint64)]
        rc = L.nflo_hwt_test(self.h, out.ctypes.data, batch, hwt, bytes(key), first_nonce, test_shrink, ctypes.byref(calls))
        assert rc == 0
        return out, calls.value

    def gaussian(self, batch, table, amplifier, key, first_nonce):
        """`batch` successive poly::set(gaussian(&prng, amplifier)) draws for the barrier table `table` (a GaussTable);
        returns (polys, number of fastrandombytes calls made)."""
        out = np.empty((batch, self.M, self.N), dtype=self.dtype)
        calls = ctypes.c_uint64()
        L = self.lib()
        L.nflo_gaussian.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_uint64,
                                    ctypes.POINTER(ctypes.c_uint64)]
        bar = np.ascontiguousarray(table.barriers)
        rc = L.nflo_gaussian(self.h, out.ctypes.data, batch, bar.ctypes.data, bar.shape[0], bar.shape[1], table.in_bytes, table.depth,
                             table.rounded_center, amplifier, bytes(key), first_nonce, ctypes.byref(calls))
        assert rc == 0, rc
        return out, calls.value

    def zo(self, batch, rho, key, first_nonce):
        out = np.empty((batch, self.M, self.N), dtype=self.dtype)
        self.lib().nflo_zo(self.h, out.ctypes.data, batch, rho, bytes(key), first_nonce)
        return out

    def spec_fwd(self, a):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        out = np.empty_like(a)
        self.lib().nflo_spec_fwd(self.h, out.ctypes.data, a.ctypes.data, a.size // (self.N * self.M))
        return out


class GaussTable:
    """A FastGaussianNoise barrier table: barriers[nb][wp] look-up words (uint8 / uint16) + the parameters that go with it."""

    def __init__(self, barriers, in_bytes, depth, rounded_center, params=None):
        self.barriers, self.in_bytes, self.depth, self.rounded_center, self.params = barriers, in_bytes, depth, rounded_center, params


def have_ref():
    return os.path.exists(REF_SO)


def ref_native_usable():
    """True when oracle/_ref/native (the reference's own -march=native -mtune=native build, CMakeCompilers.txt:17-24) was made
    on a CPU whose feature flags this host has too, i.e. it is the stated build for this host and cannot hit an illegal instruction."""
    try:
        with open(os.path.join(REF_NATIVE_DIR, "cpu_flags.txt")) as f:
            built = set(f.read().split(":", 1)[1].split())
        with open("/proc/cpuinfo") as f:
            here = next(set(line.split(":", 1)[1].split()) for line in f if line.startswith("flags"))
    except (OSError, StopIteration, IndexError):
        return False
    return built <= here and os.path.exists(os.path.join(REF_NATIVE_DIR, "libnflref.so"))


def ref_dir():
    return REF_NATIVE_DIR if ref_native_usable() else os.path.dirname(REF_SO)


class Ref:
    """The unmodified reference compiled into oracle/_ref[/native]/libnflref.so (oracle/ref_harness.cpp)."""
    _lib = None
    arch = None  # "-march=native" or "-march=x86-64-v3", whichever build was loaded

    @classmethod
    def lib(cls):
        if cls._lib is None:
            native = ref_native_usable()
            cls.arch = "-march=native -mtune=native" if native else "-march=x86-64-v3 -mtune=generic"
            L = ctypes.CDLL(os.path.join(ref_dir(), "libnflref.so"))
            L.nflref_run.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t] + [ctypes.c_void_p] * 4 + [ctypes.c_size_t, ctypes.c_int]
            L.nflref_build_flags.restype = ctypes.c_char_p
            cls._lib = L
        return cls._lib

    def __init__(self, bits, N, M):
        self.bits, self.N, self.M, self.dtype = bits, N, M, DTYPES[bits]

    def supported(self):
        z = aligned((1, self.M, self.N), self.dtype)
        return self.lib().nflref_run(OPS["add"], self.bits, self.N, self.M, z.ctypes.data, z.ctypes.data, z.ctypes.data, None, 0, 1) == 0

    def run(self, op, a, b=None, c=None, threads=1, out=None):
        a = as_aligned(np.ascontiguousarray(a, dtype=self.dtype))
        b = None if b is None else as_aligned(np.ascontiguousarray(b, dtype=self.dtype))
        c = None if c is None else as_aligned(np.ascontiguousarray(c, dtype=self.dtype))
        if out is None:
            out = aligned(a.shape, self.dtype)
        rc = self.lib().nflref_run(OPS[op], self.bits, self.N, self.M, out.ctypes.data, a.ctypes.data,
                                   None if b is None else b.ctypes.data, None if c is None else c.ctypes.data,
                                   a.size // (self.N * self.M), threads)
        assert rc == 0, f"nflref_run rc={rc} for ({self.bits},{self.N},{self.M})"
        return out

    FIXED_KEY = bytes(range(1, 33))  # what the harness's randombytes stub hands to fastrandombytes.cpp

    def sample(self, kind, batch, p0=0, p1=0):
        """(first_nonce, polys): the reference's own poly::set(uniform | non_uniform(p0, p1) | ZO_dist(p0)) run `batch` times
        with the fixed Salsa20 key of the harness."""
        out = aligned((batch, self.M, self.N), self.dtype)
        n = ctypes.c_ulonglong()
        self.lib().nflref_sample.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                             ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.POINTER(ctypes.c_ulonglong)]
        rc = self.lib().nflref_sample({"uniform": 0, "non_uniform": 1, "zo": 2, "hwt": 3}[kind], self.bits, self.N, self.M, out.ctypes.data, batch, p0, p1,
                                      ctypes.byref(n))
        assert rc == 0, rc
        return n.value, out

    def uniform(self, batch):
        return self.sample("uniform", batch)

    @classmethod
    def gaussian_table(cls, sigma, security, samples, center, in_bytes, depth, bits=64):
        """The reference's own FastGaussianNoise<in_class, T, depth>(sigma, security, samples, center): (handle, GaussTable)."""
        L = cls.lib()
        L.nflref_gaussian_create.argtypes = [ctypes.c_double, ctypes.c_uint, ctypes.c_uint, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        h = L.nflref_gaussian_create(sigma, security, samples, center, in_bytes, depth, bits)
        assert h >= 0, h
        info = (ctypes.c_longlong * 7)()
        tb = ctypes.c_double()
        assert L.nflref_gaussian_info(h, info, ctypes.byref(tb)) == 0
        nb, wp = info[0], info[1]
        bar = np.zeros((nb, wp), dtype=np.uint8 if in_bytes == 1 else np.uint16)
        assert L.nflref_gaussian_barriers(h, bar.ctypes.data_as(ctypes.c_void_p)) == 0
        params = {"nb": nb, "wp": wp, "bit_precision": info[2], "flag_ctr1": info[3], "flag_ctr2": info[4], "tail_bound": tb.value}
        return h, GaussTable(bar, in_bytes, depth, info[5], params)

    def gaussian(self, handle, batch, amplifier=1):
        """(first_nonce, nonces_used, polys): `batch` successive poly::set(gaussian(&prng, amplifier)) of the reference, fixed key."""
        out = aligned((batch, self.M, self.N), self.dtype)
        first, used = ctypes.c_ulonglong(), ctypes.c_ulonglong()
        L = self.lib()
        L.nflref_gaussian_sample.argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_ulonglong,
                                             ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_ulonglong)]
        rc = L.nflref_gaussian_sample(handle, self.N, self.M, out.ctypes.data, batch, amplifier, ctypes.byref(first), ctypes.byref(used))
        assert rc == 0, rc
        return first.value, used.value, out

    def lift(self, polys, W):
        """The reference's own poly2mpz (GMP runtime of the image): uint64[batch][N][W]."""
        polys = as_aligned(np.ascontiguousarray(polys, dtype=self.dtype))
        words = np.zeros((polys.shape[0], self.N, W), np.uint64)
        L = self.lib()
        L.nflref_lift.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t]
        rc = L.nflref_lift(0, self.bits, self.N, self.M, polys.ctypes.data, words.ctypes.data, W, polys.shape[0])
        assert rc == 0, rc
        return words

    def unlift(self, words):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        out = aligned((words.shape[0], self.M, self.N), self.dtype)
        L = self.lib()
        L.nflref_lift.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t]
        rc = L.nflref_lift(1, self.bits, self.N, self.M, out.ctypes.data, words.ctypes.data, words.shape[2], words.shape[0])
        assert rc == 0, rc
        return out

    @classmethod
    def params(cls, bits, count):
        L = cls.lib()
        arrs = [np.zeros(count, np.uint64) for _ in range(4)]
        k, mm = ctypes.c_uint64(), ctypes.c_uint64()
        L.nflref_params(bits, ctypes.c_size_t(count), *[x.ctypes.data_as(ctypes.c_void_p) for x in arrs], ctypes.byref(k), ctypes.byref(mm))
        n = min(count, mm.value)
        return {"P": [int(v) for v in arrs[0][:n]], "Pn": [int(v) for v in arrs[1][:n]], "roots": [int(v) for v in arrs[2][:n]],
                "invkmax": [int(v) for v in arrs[3][:n]], "kmax": k.value, "maxmoduli": mm.value}


def lift_words_per_coeff(P):
    q = 1
    for p in P:
        q *= int(p)
    return (q.bit_length() + 63) // 64


def crt_lift(polys, P):
    """Big-integer restatement of poly::GMP::poly2mpz (gmp.hpp:183-209): for every coefficient the unique x in [0, prod P) with
    x = a_cm (mod p_cm); returned as uint64[batch][N][W] little-endian words (the layout mpz_export(..., -1, 8, 0, 0, x) gives)."""
    P = [int(p) for p in P]
    Q = 1
    for p in P:
        Q *= p
    L = [(Q // p) * pow(Q // p, -1, p) for p in P]  # lifting_integers, gmp.hpp:137-150
    W = (Q.bit_length() + 63) // 64
    batch, M, N = polys.shape
    out = np.zeros((batch, N, W), np.uint64)
    for b in range(batch):
        cols = [polys[b, cm].astype(object) * L[cm] for cm in range(M)]
        x = cols[0]
        for c in cols[1:]:
            x = x + c
        x = x % Q
        for k in range(W):
            out[b, :, k] = ((x >> (64 * k)) & 0xFFFFFFFFFFFFFFFF).astype(np.uint64)
    return out


def crt_unlift(words, P, dtype):
    """poly::GMP::mpz2poly (gmp.hpp:211-219): a[cm][i] = x_i mod p_cm."""
    batch, N, W = words.shape
    out = np.zeros((batch, len(P), N), dtype)
    for b in range(batch):
        x = np.zeros(N, dtype=object)
        for k in range(W):
            x = x + (words[b, :, k].astype(object) << (64 * k))
        for cm, p in enumerate(P):
            out[b, cm] = (x % int(p)).astype(dtype)
    return out


def splitmix64(seed, n):
    """Deterministic uint64 stream (splitmix64), vectorised."""
    idx = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def random_polys(bits, N, M, batch, seed, P=None):
    """[batch][M][N] coefficients i.i.d. uniform in [0, p_cm)."""
    P = P if P is not None else golden_params(bits)["P"]
    r = splitmix64(seed, batch * M * N).reshape(batch, M, N)
    out = np.empty((batch, M, N), dtype=DTYPES[bits])
    for cm in range(M):
        out[:, cm, :] = (r[:, cm, :] % np.uint64(P[cm])).astype(DTYPES[bits])
    return out

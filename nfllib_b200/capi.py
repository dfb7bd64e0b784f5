"""ctypes binding of include/nflgpu.h (one-to-one; see the header for the reference function each entry point
replaces).  Device buffers are raw device pointers (ints), e.g. torch tensors' .data_ptr()."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DTYPES = {16: np.uint16, 32: np.uint32, 64: np.uint64}
HOST_OPS = {"fwd": 0, "inv": 1, "mul": 2, "mul_shoup": 3, "compute_shoup": 4, "add": 5, "sub": 6, "polymul": 8, "muladd": 9}


class NflGpuError(RuntimeError):
    pass


def lib_path():
    return os.path.join(_HERE, "libnflgpu.so")


_lib = None


def lib():
    """Loads libnflgpu.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise NflGpuError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C nfllib_b200/csrc). There is no CPU fallback.")
        L = ctypes.CDLL(path)
        vp, sz, u64p, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int
        L.nflgpu_last_error.restype = ctypes.c_char_p
        L.nflgpu_ctx_create.argtypes = [ctypes.POINTER(vp), ci, sz, sz, sz, ci, vp, vp]
        L.nflgpu_ctx_destroy.argtypes = [vp]
        L.nflgpu_ctx_info.argtypes = [vp, ctypes.POINTER(ci), ctypes.POINTER(sz), ctypes.POINTER(sz), ctypes.POINTER(ci)]
        L.nflgpu_ctx_moduli.argtypes = [vp, vp]
        L.nflgpu_ctx_launch_count.argtypes = [vp]
        L.nflgpu_ctx_launch_count.restype = ctypes.c_uint64
        L.nflgpu_params.argtypes = [ci, sz, sz, vp, vp, vp, vp]
        L.nflgpu_params_limits.argtypes = [ci, u64p, u64p, ctypes.POINTER(ctypes.c_uint)]
        L.nflgpu_batch_bytes.argtypes = [vp, sz]
        L.nflgpu_batch_bytes.restype = sz
        L.nflgpu_alloc.argtypes = [vp, sz, ctypes.POINTER(vp)]
        L.nflgpu_free.argtypes = [vp, vp]
        L.nflgpu_upload.argtypes = [vp, vp, vp, sz, vp]
        L.nflgpu_download.argtypes = [vp, vp, vp, sz, vp]
        L.nflgpu_sync.argtypes = [vp, vp]
        for name in ("nflgpu_ntt_fwd", "nflgpu_ntt_inv", "nflgpu_compute_shoup", "nflgpu_ntt_raw_fwd", "nflgpu_ntt_raw_inv"):
            getattr(L, name).argtypes = [vp, vp, vp, sz, vp]
        for name in ("nflgpu_mul", "nflgpu_add", "nflgpu_sub", "nflgpu_polymul"):
            getattr(L, name).argtypes = [vp, vp, vp, vp, sz, vp]
        for name in ("nflgpu_mul_shoup", "nflgpu_muladd"):
            getattr(L, name).argtypes = [vp, vp, vp, vp, vp, sz, vp]
        L.nflgpu_muladd_shoup.argtypes = [vp, vp, vp, vp, vp, vp, sz, vp]
        L.nflgpu_host_op.argtypes = [vp, ci, vp, vp, vp, vp, sz]
        L.nflgpu_host_op_async.argtypes = [vp, ci, vp, vp, vp, vp, sz]
        L.nflgpu_host_sync.argtypes = [vp]
        L.nflgpu_uniform.argtypes = [vp, vp, sz, ctypes.c_char_p, ctypes.c_uint64, vp]
        L.nflgpu_non_uniform.argtypes = [vp, vp, sz, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_uint64, vp]
        L.nflgpu_zo.argtypes = [vp, vp, sz, ctypes.c_uint8, ctypes.c_char_p, ctypes.c_uint64, vp]
        L.nflgpu_hwt.argtypes = [vp, vp, sz, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint64, vp]
        L.nflgpu_hwt_count.argtypes = [vp, vp, sz, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint64, u64p, vp]
        i64p = ctypes.POINTER(ctypes.c_int64)
        L.nflgpu_gaussian_create.argtypes = [ctypes.POINTER(vp), vp, ctypes.c_double, ctypes.c_uint, ctypes.c_uint, ctypes.c_double, ci, ci]
        L.nflgpu_gaussian_create_from_barriers.argtypes = [ctypes.POINTER(vp), vp, vp, sz, sz, ci, ci, ctypes.c_int64]
        L.nflgpu_gaussian_destroy.argtypes = [vp]
        L.nflgpu_gaussian_info.argtypes = [vp, i64p, ctypes.POINTER(ctypes.c_double)]
        L.nflgpu_gaussian_barriers.argtypes = [vp, vp]
        L.nflgpu_gaussian_table.argtypes = [ctypes.c_double, ctypes.c_uint, ctypes.c_uint, ctypes.c_double, ci, ci, i64p,
                                            ctypes.POINTER(ctypes.c_double), vp, sz]
        L.nflgpu_gaussian_sample.argtypes = [vp, vp, vp, sz, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_uint64, u64p, vp]
        L.nflgpu_lift_words.argtypes = [vp, ctypes.POINTER(sz)]
        L.nflgpu_poly2mpz.argtypes = [vp, vp, vp, sz, vp]
        L.nflgpu_mpz2poly.argtypes = [vp, vp, vp, sz, vp]
        L.nflgpu_eval.argtypes = [vp, vp, ctypes.POINTER(vp), sz, ctypes.c_char_p, sz, sz, vp]
        L.nflgpu_host_register.argtypes = [vp, vp, sz]
        L.nflgpu_host_unregister.argtypes = [vp, vp]
        L.nflgpu_scratch_alloc.argtypes = [vp, sz, ctypes.POINTER(vp), vp]
        L.nflgpu_scratch_free.argtypes = [vp, vp, vp]
        L.nflgpu_ctx_trim.argtypes = [vp]
        L.nflgpu_any_eq.argtypes = [vp, vp, vp, vp, sz, vp]
        L.nflgpu_any_neq.argtypes = [vp, vp, vp, vp, sz, vp]
        L.nflgpu_ipc_export.argtypes = [vp, vp, vp]
        L.nflgpu_ipc_open.argtypes = [vp, vp, ctypes.POINTER(vp)]
        L.nflgpu_ipc_close.argtypes = [vp, vp]
        L.nflgpu_gather_residues.argtypes = [vp, vp, ctypes.POINTER(vp), ctypes.POINTER(sz), ctypes.POINTER(sz), sz, sz, vp]
        L.nflgpu_poly2mpz_slabs.argtypes = [vp, vp, ctypes.POINTER(vp), ctypes.POINTER(sz), ctypes.POINTER(sz), sz, sz, vp]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise NflGpuError(f"nflgpu error {rc}: {lib().nflgpu_last_error().decode()}")


def params_limits(bits):
    k, m, b = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint()
    _check(lib().nflgpu_params_limits(bits, ctypes.byref(k), ctypes.byref(m), ctypes.byref(b)))
    return {"kmax": k.value, "maxmoduli": m.value, "modulus_bits": b.value}


def params(bits, first, count):
    arrs = [np.zeros(count, np.uint64) for _ in range(4)]
    _check(lib().nflgpu_params(bits, first, count, *[a.ctypes.data for a in arrs]))
    return dict(zip(("P", "Pn", "roots", "invkmax"), arrs))


INFO_KEYS = ("nb", "wp", "bit_precision", "flag_ctr1", "flag_ctr2", "rounded_center", "lu_size")


def gaussian_table(sigma, security, samples, center=0.0, in_bytes=1, lu_depth=2):
    """nflgpu_gaussian_table (host only): (params dict, barriers[nb][wp]) of FastGaussianNoise<in_class, T, lu_depth>(sigma,
    security, samples, center)."""
    info = (ctypes.c_int64 * 7)()
    tb = ctypes.c_double()
    _check(lib().nflgpu_gaussian_table(sigma, security, samples, center, in_bytes, lu_depth, info, ctypes.byref(tb), None, 0))
    bar = np.zeros((info[0], info[1]), dtype=np.uint8 if in_bytes == 1 else np.uint16)
    _check(lib().nflgpu_gaussian_table(sigma, security, samples, center, in_bytes, lu_depth, info, ctypes.byref(tb), bar.ctypes.data, bar.nbytes))
    d = dict(zip(INFO_KEYS, list(info)))
    d["tail_bound"] = tb.value
    return d, bar


class Gaussian:
    """nflgpu_gaussian: nfl::FastGaussianNoise<in_class, T, lu_depth> on a context's device."""

    def __init__(self, ctx, sigma=None, security=128, samples=1 << 14, center=0.0, in_bytes=1, lu_depth=2, barriers=None, rounded_center=0):
        self.ctx, self.in_bytes, self.lu_depth = ctx, in_bytes, lu_depth
        h = ctypes.c_void_p()
        if barriers is not None:
            b = np.ascontiguousarray(barriers)
            _check(lib().nflgpu_gaussian_create_from_barriers(ctypes.byref(h), ctx.h, b.ctypes.data, b.shape[0], b.shape[1], in_bytes, lu_depth,
                                                              rounded_center))
        else:
            _check(lib().nflgpu_gaussian_create(ctypes.byref(h), ctx.h, sigma, security, samples, center, in_bytes, lu_depth))
        self.h = h

    def info(self):
        info = (ctypes.c_int64 * 7)()
        tb = ctypes.c_double()
        _check(lib().nflgpu_gaussian_info(self.h, info, ctypes.byref(tb)))
        d = dict(zip(INFO_KEYS, list(info)))
        d["tail_bound"] = tb.value
        return d

    def sample(self, dst, batch, key, first_nonce, amplifier=1, stream=0):
        """nflgpu_gaussian_sample; returns the number of nonces (fastrandombytes calls) the batch consumed."""
        used = ctypes.c_uint64()
        _check(lib().nflgpu_gaussian_sample(self.ctx.h, self.h, dst, batch, amplifier, bytes(key), first_nonce, ctypes.byref(used), stream))
        return used.value

    def close(self):
        if getattr(self, "h", None):
            lib().nflgpu_gaussian_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class Context:
    """nflgpu_ctx: the per-(limb, degree, nmoduli) state of nfl::poly<T,Degree,NbModuli>::core on one device."""

    def __init__(self, bits, degree, nmoduli, device=0, first_modulus=0, moduli=None, roots=None):
        self.bits, self.degree, self.nmoduli, self.device = bits, degree, nmoduli, device
        self.dtype = DTYPES[bits]
        h = ctypes.c_void_p()
        keep = None
        pm = pr = None
        if moduli is not None:
            keep = (np.ascontiguousarray(moduli, dtype=np.uint64), np.ascontiguousarray(roots, dtype=np.uint64))
            pm, pr = keep[0].ctypes.data, keep[1].ctypes.data
        _check(lib().nflgpu_ctx_create(ctypes.byref(h), bits, degree, nmoduli, first_modulus, device, pm, pr))
        self.h = h
        m = np.zeros(nmoduli, np.uint64)
        _check(lib().nflgpu_ctx_moduli(self.h, m.ctypes.data))
        self.moduli = m

    def close(self):
        if getattr(self, "h", None):
            lib().nflgpu_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 -- interpreter shutdown: the module globals may already be gone
            pass

    def batch_bytes(self, batch):
        return lib().nflgpu_batch_bytes(self.h, batch)

    @property
    def launch_count(self):
        return lib().nflgpu_ctx_launch_count(self.h)

    # ---- raw device-pointer calls (stream = cudaStream_t handle as int, 0 = default stream) ----
    def alloc(self, batch):
        p = ctypes.c_void_p()
        _check(lib().nflgpu_alloc(self.h, batch, ctypes.byref(p)))
        return p.value

    def free(self, dptr):
        _check(lib().nflgpu_free(self.h, dptr))

    def upload(self, dptr, host, batch, stream=0):
        _check(lib().nflgpu_upload(self.h, dptr, host.ctypes.data, batch, stream))

    def download(self, host, dptr, batch, stream=0):
        _check(lib().nflgpu_download(self.h, host.ctypes.data, dptr, batch, stream))

    def sync(self, stream=0):
        _check(lib().nflgpu_sync(self.h, stream))

    def ntt_fwd(self, dst, src, batch, stream=0):
        _check(lib().nflgpu_ntt_fwd(self.h, dst, src, batch, stream))

    def ntt_inv(self, dst, src, batch, stream=0):
        _check(lib().nflgpu_ntt_inv(self.h, dst, src, batch, stream))

    def ntt_raw_fwd(self, dst, src, batch, stream=0):
        _check(lib().nflgpu_ntt_raw_fwd(self.h, dst, src, batch, stream))

    def ntt_raw_inv(self, dst, src, batch, stream=0):
        _check(lib().nflgpu_ntt_raw_inv(self.h, dst, src, batch, stream))

    def mul(self, dst, a, b, batch, stream=0):
        _check(lib().nflgpu_mul(self.h, dst, a, b, batch, stream))

    def add(self, dst, a, b, batch, stream=0):
        _check(lib().nflgpu_add(self.h, dst, a, b, batch, stream))

    def sub(self, dst, a, b, batch, stream=0):
        _check(lib().nflgpu_sub(self.h, dst, a, b, batch, stream))

    def mul_shoup(self, dst, a, b, bprime, batch, stream=0):
        _check(lib().nflgpu_mul_shoup(self.h, dst, a, b, bprime, batch, stream))

    def compute_shoup(self, dst, a, batch, stream=0):
        _check(lib().nflgpu_compute_shoup(self.h, dst, a, batch, stream))

    def muladd(self, dst, a, b, c, batch, stream=0):
        _check(lib().nflgpu_muladd(self.h, dst, a, b, c, batch, stream))

    def muladd_shoup(self, dst, a, b, c, cprime, batch, stream=0):
        _check(lib().nflgpu_muladd_shoup(self.h, dst, a, b, c, cprime, batch, stream))

    def uniform(self, dst, batch, key, first_nonce, stream=0):
        """nflgpu_uniform: `batch` poly::set(uniform) draws from the Salsa20 stream (key, first_nonce + i)."""
        _check(lib().nflgpu_uniform(self.h, dst, batch, bytes(key), first_nonce, stream))

    def non_uniform(self, dst, batch, upper_bound, amplifier, key, first_nonce, stream=0):
        _check(lib().nflgpu_non_uniform(self.h, dst, batch, upper_bound, amplifier, bytes(key), first_nonce, stream))

    def hwt(self, dst, batch, hwt, key, first_nonce, stream=0):
        _check(lib().nflgpu_hwt(self.h, dst, batch, hwt, bytes(key), first_nonce, stream))

    def hwt_count(self, dst, batch, hwt, key, first_nonce, stream=0):
        """nflgpu_hwt_count: like hwt(), returns the number of nonces the batch consumed (synchronises the stream)."""
        used = ctypes.c_uint64()
        _check(lib().nflgpu_hwt_count(self.h, dst, batch, hwt, bytes(key), first_nonce, ctypes.byref(used), stream))
        return used.value

    def zo(self, dst, batch, rho, key, first_nonce, stream=0):
        _check(lib().nflgpu_zo(self.h, dst, batch, rho, bytes(key), first_nonce, stream))

    def lift_words(self):
        w = ctypes.c_size_t()
        _check(lib().nflgpu_lift_words(self.h, ctypes.byref(w)))
        return w.value

    def poly2mpz(self, dst_words, src_polys, batch, stream=0):
        _check(lib().nflgpu_poly2mpz(self.h, dst_words, src_polys, batch, stream))

    def mpz2poly(self, dst_polys, src_words, batch, stream=0):
        _check(lib().nflgpu_mpz2poly(self.h, dst_polys, src_words, batch, stream))

    def eval(self, dst, operands, program, batch, stream=0):
        """nflgpu_eval: `operands` = list of device pointers, `program` = postfix bytes (see include/nflgpu.h)."""
        arr = (ctypes.c_void_p * len(operands))(*operands)
        prog = bytes(program)
        _check(lib().nflgpu_eval(self.h, dst, arr, len(operands), prog, len(prog), batch, stream))

    def polymul(self, dst, a, b, batch, stream=0):
        _check(lib().nflgpu_polymul(self.h, dst, a, b, batch, stream))

    def scratch_alloc(self, batch, stream=0):
        p = ctypes.c_void_p()
        _check(lib().nflgpu_scratch_alloc(self.h, batch, ctypes.byref(p), stream))
        return p.value

    def scratch_free(self, dptr, stream=0):
        _check(lib().nflgpu_scratch_free(self.h, dptr, stream))

    def trim(self):
        _check(lib().nflgpu_ctx_trim(self.h))

    def any_eq(self, flags, a, b, batch, stream=0):
        """nflgpu_any_eq: flags[i] (uint8, device) = any coefficient of a[i] equals that of b[i] (ops.hpp:81-117)."""
        _check(lib().nflgpu_any_eq(self.h, flags, a, b, batch, stream))

    def any_neq(self, flags, a, b, batch, stream=0):
        _check(lib().nflgpu_any_neq(self.h, flags, a, b, batch, stream))

    # ---- residues sharded over devices ----
    def ipc_export(self, dptr):
        """64 opaque bytes naming the allocation `dptr` (from alloc()) for another process on this node."""
        h = ctypes.create_string_buffer(64)
        _check(lib().nflgpu_ipc_export(self.h, dptr, h))
        return h.raw

    def ipc_open(self, handle):
        h = ctypes.create_string_buffer(bytes(handle), 64)
        p = ctypes.c_void_p()
        _check(lib().nflgpu_ipc_open(self.h, h, ctypes.byref(p)))
        return p.value

    def ipc_close(self, peer_ptr):
        _check(lib().nflgpu_ipc_close(self.h, peer_ptr))

    def gather_residues(self, dst_full, slabs, batch, stream=0):
        """nflgpu_gather_residues on this (full) context: slabs = [(device pointer, first_residue, nresidues), ...]."""
        n = len(slabs)
        ptrs = (ctypes.c_void_p * n)(*[s[0] for s in slabs])
        first = (ctypes.c_size_t * n)(*[s[1] for s in slabs])
        cnt = (ctypes.c_size_t * n)(*[s[2] for s in slabs])
        _check(lib().nflgpu_gather_residues(self.h, dst_full, ptrs, first, cnt, n, batch, stream))

    # ---- host-buffer call (numpy in, numpy out): H2D + kernel(s) + D2H inside the library ----
    def host_op(self, op, a, b=None, c=None, out=None, wait=True):
        """nflgpu_host_op; wait=False -> nflgpu_host_op_async (operands and `out` must stay alive and untouched until host_sync())."""
        a = np.ascontiguousarray(a, dtype=self.dtype)
        batch = a.size // (self.degree * self.nmoduli)
        if out is None:
            out = np.empty_like(a)
        ops = [None if x is None else np.ascontiguousarray(x, dtype=self.dtype) for x in (b, c)]
        fn = lib().nflgpu_host_op if wait else lib().nflgpu_host_op_async
        if not wait:
            self._inflight = getattr(self, "_inflight", []) + [(a, ops, out)]  # keep the arrays alive until host_sync
        _check(fn(self.h, HOST_OPS[op], out.ctypes.data, a.ctypes.data,
                  None if ops[0] is None else ops[0].ctypes.data,
                  None if ops[1] is None else ops[1].ctypes.data, batch))
        if wait:
            self._inflight = []
        return out

    def host_sync(self):
        """nflgpu_host_sync: every earlier host_op(..., wait=False) is complete when this returns."""
        try:
            _check(lib().nflgpu_host_sync(self.h))
        finally:
            self._inflight = []

    def poly2mpz_slabs(self, dst_words, slabs, batch, stream=0):
        """nflgpu_poly2mpz_slabs: the CRT lift reading each residue from the slab (local or peer) that holds it."""
        n = len(slabs)
        ptrs = (ctypes.c_void_p * n)(*[s[0] for s in slabs])
        first = (ctypes.c_size_t * n)(*[s[1] for s in slabs])
        cnt = (ctypes.c_size_t * n)(*[s[2] for s in slabs])
        _check(lib().nflgpu_poly2mpz_slabs(self.h, dst_words, ptrs, first, cnt, n, batch, stream))

    def host_register(self, array):
        """nflgpu_host_register: page-lock a numpy array so that host_op DMAs it directly."""
        _check(lib().nflgpu_host_register(self.h, array.ctypes.data, array.nbytes))

    def host_unregister(self, array):
        _check(lib().nflgpu_host_unregister(self.h, array.ctypes.data))

    # ---- convenience for tests: run a device-resident op on numpy data through alloc/upload/.../download ----
    def run_device(self, op, a, b=None, c=None, d=None, inplace=False):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        batch = a.size // (self.degree * self.nmoduli)
        bufs = []
        try:
            for x in (a, b, c, d):
                if x is None:
                    bufs.append(None)
                    continue
                x = np.ascontiguousarray(x, dtype=self.dtype)
                p = self.alloc(batch)
                self.upload(p, x, batch)
                bufs.append(p)
            out = bufs[0] if inplace else self.alloc(batch)
            args = [out] + [p for p in bufs if p is not None]
            getattr(self, op)(*args, batch)
            host = np.empty_like(a)
            self.download(host, out, batch)
            self.sync()
            if not inplace:
                self.free(out)
            return host
        finally:
            for p in bufs:
                if p is not None:
                    self.free(p)

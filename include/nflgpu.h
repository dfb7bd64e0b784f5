/*
 * nflgpu.h — C ABI of the B200-native NTT / pointwise hot path of NFLlib.
 *
 * This is the drop-in boundary.  The reference has no FFI layer: its hot path sits
 * behind the compile-time template surface nfl::poly<T,Degree,NbModuli> (include/nfl/poly.hpp:82-309) whose
 * backend seam is the CC_SIMD tag (include/nfl/arch.hpp:6-18).  The C++11 header include/nfl_b200.hpp keeps
 * that template surface and forwards to the entry points below; each entry point names the reference
 * function it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross this boundary;
 *   - every function returns 0 on success or a negative nflgpu_status; nflgpu_last_error() returns a
 *     thread-local message for the last failure on the calling thread; nothing throws across the ABI;
 *   - a "batch buffer" is `batch` polynomials laid out exactly like an array of nfl::poly
 *     (poly.hpp:87-88,156-157):  limb[batch][nmoduli][degree], residue-major, little-endian limbs of
 *     `limb_bits` bits; device buffers must be 16-byte aligned;
 *   - coefficients must be canonical, i.e. < p_cm (the reference's contract, ops.hpp:131,148,190); outputs
 *     are canonical and bit-identical to the reference's;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream); calls on one stream are
 *     ordered; a context is bound to one device; distinct contexts are independent;
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point fails with
 *     NFLGPU_ERR_CUDA.
 */
#ifndef NFLGPU_H
#define NFLGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nflgpu_ctx nflgpu_ctx;

typedef enum {
  NFLGPU_OK = 0,
  NFLGPU_ERR_ARG = -1,         /* bad argument (size, alignment, null pointer) */
  NFLGPU_ERR_UNSUPPORTED = -2, /* (limb_bits, degree, nmoduli) outside what the kernels are built for */
  NFLGPU_ERR_CUDA = -3,        /* CUDA runtime / driver error, or no device */
  NFLGPU_ERR_ALLOC = -4        /* host or device allocation failed */
} nflgpu_status;

/* ---- context ------------------------------------------------------------------------------------------ */

/* Creates the per-(limb, degree, nmoduli) state that nfl::poly<T,Degree,NbModuli>::core::initialize()
 * builds at static-init time (core.hpp:625-686): moduli, the psi / psi^-1 twiddle tables with their Shoup
 * companions, N^-1.  Tables are derived from the moduli and the primitive 2*kMaxPolyDegree-th roots exactly
 * as core.hpp:640-665 derives phi and N^-1, then uploaded to `device`.
 *   limb_bits      16, 32 or 64  (params<uint16_t|uint32_t|uint64_t>, params.hpp:12,44,83)
 *   degree         power of two, 32 bytes <= degree*limb_bits/8, degree <= params<T>::kMaxPolyDegree (512 / 32768 / 2^20)
 *   first_modulus  index of the first modulus in NFLlib's table; a context covers
 *                  P[first_modulus .. first_modulus+nmoduli) — nonzero when residues are sharded over GPUs
 *   moduli, roots  optional caller-provided tables of `nmoduli` uint64_t each (params<T>::P and
 *                  params<T>::primitive_roots, widened); pass NULL for both to use the built-in derivation
 *                  of NFLlib's tables (nflgpu_params_*). */
int nflgpu_ctx_create(nflgpu_ctx **ctx, int limb_bits, size_t degree, size_t nmoduli, size_t first_modulus,
                      int device, const uint64_t *moduli, const uint64_t *roots);
int nflgpu_ctx_destroy(nflgpu_ctx *ctx);

const char *nflgpu_last_error(void);

/* Introspection. */
int nflgpu_ctx_info(const nflgpu_ctx *ctx, int *limb_bits, size_t *degree, size_t *nmoduli, int *device);
/* Copies the context's moduli (uint64_t[nmoduli]) — nfl::poly::get_modulus(n), poly.hpp:163. */
int nflgpu_ctx_moduli(const nflgpu_ctx *ctx, uint64_t *out);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t nflgpu_ctx_launch_count(const nflgpu_ctx *ctx);

/* ---- NFLlib parameter tables (params.hpp:12-119; lib/params/params.cpp) -------------------------------- *
 * Re-derived, not copied: P[i] are the primes 2^(w-2) - k*2*kMax + 1 in descending order, Pn[i] =
 * floor(2^(2w)/P[i]) - 2^(w+2), roots[i] = g^((P[i]-1)/(2*kMax)) for the least primitive root g of P[i],
 * invkmax[i] = kMax^-1 mod P[i].  Any output pointer may be NULL.  `first + count` must not exceed
 * params<T>::kMaxNbModuli (2 / 291 / 1000). */
int nflgpu_params(int limb_bits, size_t first, size_t count, uint64_t *P, uint64_t *Pn, uint64_t *roots,
                  uint64_t *invkmax);
int nflgpu_params_limits(int limb_bits, uint64_t *kMaxPolyDegree, uint64_t *kMaxNbModuli,
                         unsigned *kModulusBitsize);

/* ---- device batch buffers ------------------------------------------------------------------------------ */

size_t nflgpu_batch_bytes(const nflgpu_ctx *ctx, size_t batch);
int nflgpu_alloc(nflgpu_ctx *ctx, size_t batch, void **dptr);
int nflgpu_free(nflgpu_ctx *ctx, void *dptr);
/* Stream-ordered temporaries from a pool owned by the context (cudaMallocFromPoolAsync / cudaFreeAsync): what the C++ header
 * uses for the operands of a single-poly expression instead of a cudaMalloc / cudaFree pair per leaf.  Freed blocks stay
 * cached in the pool; nflgpu_ctx_trim returns them to the driver (it synchronises the device), nflgpu_ctx_destroy drops the pool. */
int nflgpu_scratch_alloc(nflgpu_ctx *ctx, size_t batch, void **dptr, void *stream);
int nflgpu_scratch_free(nflgpu_ctx *ctx, void *dptr, void *stream);
int nflgpu_ctx_trim(nflgpu_ctx *ctx);
int nflgpu_upload(nflgpu_ctx *ctx, void *dst_dev, const void *src_host, size_t batch, void *stream);
int nflgpu_download(nflgpu_ctx *ctx, void *dst_host, const void *src_dev, size_t batch, void *stream);
int nflgpu_sync(nflgpu_ctx *ctx, void *stream);

/* ---- transforms on device-resident batches -------------------------------------------------------------- */

/* nfl::poly::ntt_pow_phi() (poly.hpp:167 -> core.hpp:594-600 -> core.hpp:455-532 + algos.hpp:16-73):
 * dst[b][cm][j] = sum_i src[b][cm][i] * phi_cm^(i*(2*bitrev(j)+1)) mod p_cm, canonical.  dst may equal src. */
int nflgpu_ntt_fwd(nflgpu_ctx *ctx, void *dst, const void *src, size_t batch, void *stream);
/* nfl::poly::invntt_pow_invphi() (poly.hpp:168 -> core.hpp:608-614, 539-557, permut.hpp): exact inverse of
 * nflgpu_ntt_fwd (bit-reversed input order, natural output order, canonical).  dst may equal src. */
int nflgpu_ntt_inv(nflgpu_ctx *ctx, void *dst, const void *src, size_t batch, void *stream);

/* The cyclic transform underneath, without the phi twist: poly::core::ntt(x, wtab, winvtab, p) (core.hpp:455-532; the
 * function tests/ntt_perfs.cpp:122-134,165-171 times through its friend proxy) on every residue:
 * dst[b][cm][j] = sum_i src[b][cm][i] * omega_cm^(i*bitrev(j)) mod p_cm, canonical.  Tables are built on first use. */
int nflgpu_ntt_raw_fwd(nflgpu_ctx *ctx, void *dst, const void *src, size_t batch, void *stream);
/* poly::core::inv_ntt (core.hpp:539-557: bit-reverse, core::ntt with omega^-1, bit-reverse): the unscaled inverse,
 * nflgpu_ntt_raw_inv(nflgpu_ntt_raw_fwd(x)) = degree * x mod p (the reference folds N^-1 into the twist that follows). */
int nflgpu_ntt_raw_inv(nflgpu_ctx *ctx, void *dst, const void *src, size_t batch, void *stream);

/* ---- pointwise functors (the expression evaluator core.hpp:24-37 applied to one functor) ---------------- */

/* operator*  = ops::mulmod        (ops.hpp:184-219) */
int nflgpu_mul(nflgpu_ctx *ctx, void *dst, const void *a, const void *b, size_t batch, void *stream);
/* operator+  = ops::addmod        (ops.hpp:124-135) */
int nflgpu_add(nflgpu_ctx *ctx, void *dst, const void *a, const void *b, size_t batch, void *stream);
/* operator-  = ops::submod        (ops.hpp:141-151) */
int nflgpu_sub(nflgpu_ctx *ctx, void *dst, const void *a, const void *b, size_t batch, void *stream);
/* shoup(a*b, bprime) = ops::mulmod_shoup (ops.hpp:225-242, rewrite rule ops.hpp:266-277) */
int nflgpu_mul_shoup(nflgpu_ctx *ctx, void *dst, const void *a, const void *b, const void *bprime,
                     size_t batch, void *stream);
/* compute_shoup(a) = ops::compute_shoup (ops.hpp:165-177): floor((a mod p) * 2^w / p) */
int nflgpu_compute_shoup(nflgpu_ctx *ctx, void *dst, const void *a, size_t batch, void *stream);
/* a + b*c in one pass = the fused expression `a + b*c` (tests/nfllib_demo_main_op.cpp:232-258);
 * same values as ops::muladd (opt/ops.hpp:9-48). */
int nflgpu_muladd(nflgpu_ctx *ctx, void *dst, const void *a, const void *b, const void *c, size_t batch,
                  void *stream);
/* a + shoup(b*c, cprime) = ops::muladd_shoup (opt/ops.hpp:56-78), canonical result. */
int nflgpu_muladd_shoup(nflgpu_ctx *ctx, void *dst, const void *a, const void *b, const void *c,
                        const void *cprime, size_t batch, void *stream);

/* operator== / operator!= on device-resident batches, with the reference's semantics (expr::operator bool over
 * ops::eqmod / ops::neqmod, ops.hpp:81-117): flags[b] = 1 iff ANY coefficient of a[b] equals (any_eq) / differs from (any_neq)
 * the same coefficient of b[b], else 0.  `flags` is uint8_t[batch] in device memory. */
int nflgpu_any_eq(nflgpu_ctx *ctx, uint8_t *flags, const void *a, const void *b, size_t batch, void *stream);
int nflgpu_any_neq(nflgpu_ctx *ctx, uint8_t *flags, const void *a, const void *b, size_t batch, void *stream);

/* Fused evaluation of an ARBITRARY expression tree in one pass over memory — what the reference's expression
 * templates do (ops::expr ops.hpp:52-97, _make_op rewrite ops.hpp:249-277, evaluator core.hpp:24-37): `c = a + b*d - e`
 * reads a, b, d, e once and writes c once.  `program` is the tree in postfix order, one byte per token:
 *     0x00..0x07  push operands[k]
 *     0x10 add   0x11 sub   0x12 mul            pop y, pop x, push x (op) y         (addmod / submod / mulmod)
 *     0x13 mul_shoup                            pop y', pop y, pop x, push x*y       (mulmod_shoup with y' = shoup(y))
 *     0x14 compute_shoup                        pop x, push floor(x * 2^w / p)
 * Limits: at most 8 operands, 32 tokens, stack depth 8; the program must leave exactly one value.  dst may alias an
 * operand. */
int nflgpu_eval(nflgpu_ctx *ctx, void *dst, const void *const *operands, size_t noperands, const uint8_t *program,
                size_t ntokens, size_t batch, void *stream);

/* ---- fused negacyclic product ---------------------------------------------------------------------------- */

/* dst = invntt_pow_invphi( ntt_pow_phi(a) * ntt_pow_phi(b) )  — the four reference calls of
 * tests/nfllib_demo_main_op.cpp:31-45 — i.e. a*b mod (X^N + 1, p_cm), coefficient domain in and out. */
int nflgpu_polymul(nflgpu_ctx *ctx, void *dst, const void *a, const void *b, size_t batch, void *stream);

/* ---- sampler on the input side of the path --------------------------------------------------------------- */

/* `batch` successive poly::set(nfl::uniform()) draws (core.hpp:150-187) written straight into a device batch: polynomial
 * i is filled from the Salsa20/20 keystream (key, nonce = first_nonce + i) that nfl::fastrandombytes
 * (lib/prng/fastrandombytes.cpp:21-34) would produce for it, every limb masked to its modulus' bit length and reduced
 * by one conditional subtraction.  With the same 32-byte key the result is bit-identical to the reference's draws; the
 * caller owns key management (the reference keys itself once from /dev/urandom). */
int nflgpu_uniform(nflgpu_ctx *ctx, void *dst, size_t batch, const uint8_t key[32], uint64_t first_nonce, void *stream);
/* poly::set(nfl::non_uniform(upper_bound, amplifier)) (core.hpp:190-278): centred noise in (-upper_bound, upper_bound)
 * times `amplifier`, the same value in every residue (negative values stored as p_cm - |v|); one keystream of `degree`
 * limbs per polynomial.  upper_bound >= a modulus is rejected like the reference's std::runtime_error (core.hpp:201-206). */
int nflgpu_non_uniform(nflgpu_ctx *ctx, void *dst, size_t batch, uint64_t upper_bound, uint64_t amplifier,
                       const uint8_t key[32], uint64_t first_nonce, void *stream);
/* poly::set(nfl::ZO_dist(rho)) (core.hpp:338-349, poly.hpp:59-62): coefficients in {-1, 0, 1} with P(+-1) = (rho/255)/2
 * each, from one keystream byte per coefficient; -1 / +1 are stored as p_cm - 1 / p_cm + 1 exactly as the reference does. */
int nflgpu_zo(nflgpu_ctx *ctx, void *dst, size_t batch, uint8_t rho, const uint8_t key[32], uint64_t first_nonce, void *stream);
/* poly::set(nfl::hwt_dist(hwt)) (core.hpp:355-392, poly.hpp:54-57): exactly `hwt` coefficients are +-1 (p_cm - 1 / p_cm + 1),
 * chosen by reservoir sampling with rejection-sampled indices.  A polynomial normally consumes ceil((degree - hwt) / hwt) + 1
 * fastrandombytes calls (nonces); an index rejection (probability < 2^-44 per draw) can cost one more, which moves the start
 * of every later polynomial in the reference's sequential stream.  The device draws all polynomials in parallel from the
 * normal starts, then a repair kernel (idle unless that happened) re-draws what a longer polynomial displaced: the result is
 * the reference's nonce sequence in every case.  nflgpu_hwt_count also reports the nonces the batch consumed (where the stream
 * continues) and synchronises `stream` to do so; nflgpu_hwt is the asynchronous form. */
int nflgpu_hwt(nflgpu_ctx *ctx, void *dst, size_t batch, uint32_t hwt, const uint8_t key[32], uint64_t first_nonce, void *stream);
int nflgpu_hwt_count(nflgpu_ctx *ctx, void *dst, size_t batch, uint32_t hwt, const uint8_t key[32], uint64_t first_nonce,
                     uint64_t *nonces_used, void *stream);

/* poly::set(nfl::gaussian<in_class, T, lu_depth>(&prng, amplifier)) (core.hpp:284-325, poly.hpp:61-67) over
 * nfl::FastGaussianNoise<in_class, T, lu_depth>(sigma, security, samples, center) (prng/FastGaussianNoise.hpp).
 *
 * nflgpu_gaussian_create builds what the reference's constructor builds — tail bound, precision, the cumulative
 * distribution ("barrier") table in MPFR arithmetic, the look-up tables (FastGaussianNoise.hpp:233-476) — bit-identical to
 * the reference's: the table is computed with the same MPFR calls, through the MPFR/GMP runtimes (libmpfr.so.6,
 * libgmp.so.10) opened on first use; without them it returns NFLGPU_ERR_UNSUPPORTED (MPFR is a build dependency of the
 * reference itself).  nflgpu_gaussian_create_from_barriers takes a barrier table computed elsewhere instead:
 * nbarriers rows of word_precision look-up words of in_bytes bytes, most significant word first, as the reference holds them.
 *   in_bytes 1 | 2 = in_class uint8_t | uint16_t;  lu_depth 1 | 2;  supported shapes: (1, 1), (1, 2), (2, 1).
 * A sampler belongs to the context's device and can be used with any context on that device.
 *
 * nflgpu_gaussian_sample writes `batch` successive draws into a device batch, the reference's PRNG being at nonce
 * first_nonce before the first one.  getNoise() refills its keystream buffer a data-dependent number of times
 * (FastGaussianNoise.hpp:601-610), so the nonce a draw starts with depends on all draws before it; the device evaluates every
 * possible starting nonce of a window in parallel and then follows the chain (csrc/sampler.cu), which reproduces the
 * sequential reference bit for bit.  *nonces_used (optional) receives the number of fastrandombytes calls the batch made, i.e.
 * where the stream continues.  The call synchronises `stream` before returning. */
typedef struct nflgpu_gaussian nflgpu_gaussian;
int nflgpu_gaussian_create(nflgpu_gaussian **g, nflgpu_ctx *ctx, double sigma, unsigned security, unsigned samples, double center,
                           int in_bytes, int lu_depth);
int nflgpu_gaussian_create_from_barriers(nflgpu_gaussian **g, nflgpu_ctx *ctx, const void *barriers, size_t nbarriers,
                                         size_t word_precision, int in_bytes, int lu_depth, int64_t rounded_center);
int nflgpu_gaussian_destroy(nflgpu_gaussian *g);
/* The host part alone (no device needed): the parameters and the barrier table nflgpu_gaussian_create would build.
 * info as in nflgpu_gaussian_info (the flag counters are those of lu_depth); `barriers` may be NULL to query the size
 * (info[0] * info[1] * in_bytes bytes), otherwise `capacity` bytes must hold it. */
int nflgpu_gaussian_table(double sigma, unsigned security, unsigned samples, double center, int in_bytes, int lu_depth,
                          int64_t info[7], double *tail_bound, void *barriers, size_t capacity);
/* info[0..6] = number of barriers, word precision, bit precision, flagged first-level entries, flagged second-level entries,
 * rounded center, look-up table size — the private members of the reference object, for parity checks */
int nflgpu_gaussian_info(const nflgpu_gaussian *g, int64_t info[7], double *tail_bound);
/* copies the barrier table: info[0] * info[1] * in_bytes bytes */
int nflgpu_gaussian_barriers(const nflgpu_gaussian *g, void *out);
int nflgpu_gaussian_sample(nflgpu_ctx *ctx, const nflgpu_gaussian *g, void *dst, size_t batch, uint64_t amplifier,
                           const uint8_t key[32], uint64_t first_nonce, uint64_t *nonces_used, void *stream);

/* ---- CRT lift: the consumer that needs all residues of a polynomial on one device ------------------------- */

/* poly::GMP::poly2mpz / mpz2poly (include/nfl/gmp.hpp:183-219) without GMP types: a lifted coefficient is
 * W = ceil(bits(prod p_cm) / 64) little-endian 64-bit words (what mpz_export(w, 0, -1, 8, 0, 0, x) writes);
 * word buffers are uint64_t[batch][degree][W] in device memory.  nflgpu_lift_words reports W (at most 16, i.e. moduli
 * products up to 1024 bits; larger ones return NFLGPU_ERR_UNSUPPORTED).
 *   poly2mpz: x_i = the unique integer in [0, prod p_cm) with x_i = src[cm][i] (mod p_cm) for every residue
 *   mpz2poly: dst[cm][i] = x_i mod p_cm */
int nflgpu_lift_words(nflgpu_ctx *ctx, size_t *words_per_coefficient);
int nflgpu_poly2mpz(nflgpu_ctx *ctx, uint64_t *dst_words, const void *src_polys, size_t batch, void *stream);
int nflgpu_mpz2poly(nflgpu_ctx *ctx, void *dst_polys, const uint64_t *src_words, size_t batch, void *stream);

/* ---- residues sharded over GPUs: gathering the full RNS vector ----------------------- *
 * The path itself needs no exchange: a context created with first_modulus = r0 and nmoduli = k transforms the slab
 * limb[batch][k][degree] of residues r0 .. r0+k-1 on its own device.  Only a consumer that needs every residue of a
 * polynomial on one device (the CRT lift, gmp.hpp:183-209) gathers: nflgpu_gather_residues writes `nslabs` slabs
 * limb[batch][nresidues[s]][degree] into dst_full = limb[batch][nmoduli][degree] of the FULL context `ctx` at residue offset
 * first_residue[s], one strided copy-engine transfer per slab (peer memory is pulled over NVLink / NVSwitch).
 * A slab may live on this device or on a peer: one process per GPU exports its slab with nflgpu_ipc_export (a CUDA IPC
 * handle, 64 opaque bytes to ship through any channel — MPI, torch.distributed, a pipe), the gathering process maps it with
 * nflgpu_ipc_open and passes the mapped pointer.  `dptr` must be the start of an allocation made by nflgpu_alloc.
 * Ordering across processes (the slab is complete before it is read, not overwritten while it is read) is the caller's:
 * synchronise the producing stream and pass a barrier, as with any peer-memory transfer. */
typedef struct { unsigned char bytes[64]; } nflgpu_ipc_handle;
int nflgpu_ipc_export(nflgpu_ctx *ctx, const void *dptr, nflgpu_ipc_handle *out);
int nflgpu_ipc_open(nflgpu_ctx *ctx, const nflgpu_ipc_handle *handle, void **peer_ptr);
int nflgpu_ipc_close(nflgpu_ctx *ctx, void *peer_ptr);
int nflgpu_gather_residues(nflgpu_ctx *ctx, void *dst_full, const void *const *slabs, const size_t *first_residue,
                           const size_t *nresidues, size_t nslabs, size_t batch, void *stream);
/* The consumer fused with the gather: nflgpu_poly2mpz reading every residue straight from the slab that holds it (same slab
 * description as nflgpu_gather_residues; the slabs must cover each residue of the full context exactly once).  Peer slabs are
 * read over NVLink by the lift kernel itself, coalesced along the coefficient index; no gathered copy is ever written. */
int nflgpu_poly2mpz_slabs(nflgpu_ctx *ctx, uint64_t *dst_words, const void *const *slabs, const size_t *first_residue,
                          const size_t *nresidues, size_t nslabs, size_t batch, void *stream);

/* ---- host-buffer entry points (what a single host nfl::poly call maps to) -------------------------------- *
 * Same operations on HOST buffers: host->device copy, kernel(s), device->host copy, cut into chunks that move through a
 * ring of device buffers on three streams (uploads / kernels / downloads, chained by events), so both PCIe directions
 * and the kernels overlap; a small call (one chunk, <= NFLGPU_HOST_SMALL_KIB = 8 MiB per operand) lets the kernel read and write
 * mapped pinned memory directly instead (the latency path: a single polynomial, a few hundred at most).  op: 0 fwd, 1 inv, 2 mul, 3 mul_shoup, 4 compute_shoup, 5 add,
 * 6 sub, 8 polymul, 9 muladd, 10 raw_fwd (core::ntt), 11 raw_inv (core::inv_ntt).  Unused operands are NULL.  dst may be
 * one of the operands (in place).  These are the calls bench.py's e2e figure times.
 * Not re-entrant per context (the ring belongs to the context): serialise calls on one context, or use
 * one context per host thread. */
int nflgpu_host_op(nflgpu_ctx *ctx, int op, void *dst_host, const void *a_host, const void *b_host,
                   const void *c_host, size_t batch);
/* The same call without the final wait: it returns once the last chunk has been queued (it only blocks while the ring is
 * full), so a caller that streams several independent batches keeps the uploads of the next call running under the
 * downloads of this one.  Operands and destination must stay valid and untouched until nflgpu_host_sync (or any
 * nflgpu_host_op on the same context) has returned; only then is dst complete -- in particular the destination of one
 * asynchronous call cannot be an operand of another before that (the calls in flight are independent batches).  An error
 * reported by a later call or by nflgpu_host_sync may belong to an earlier asynchronous call. */
int nflgpu_host_op_async(nflgpu_ctx *ctx, int op, void *dst_host, const void *a_host, const void *b_host,
                         const void *c_host, size_t batch);
int nflgpu_host_sync(nflgpu_ctx *ctx);
/* Pageable host arrays (e.g. posix_memalign'ed nfl::poly[], tests/tools.h:6-17) go through pinned staging buffers with an extra
 * host copy each way (a few host threads, NFLGPU_HOST_COPY_THREADS).  A caller that keeps its arrays can page-lock them once
 * instead — nflgpu_host_register is cudaHostRegister without the CUDA headers — after which nflgpu_host_op DMAs them directly,
 * like memory from cudaHostAlloc.  Unregister before freeing the memory. */
int nflgpu_host_register(nflgpu_ctx *ctx, void *host_ptr, size_t bytes);
int nflgpu_host_unregister(nflgpu_ctx *ctx, void *host_ptr);

#ifdef __cplusplus
}
#endif
#endif /* NFLGPU_H */

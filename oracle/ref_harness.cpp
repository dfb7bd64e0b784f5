// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Thin extern "C" driver around the UNMODIFIED reference headers (compiled from /root/reference where
// they lie; nothing is copied).  It is built by oracle/Makefile into oracle/_ref/libnflref.so and is used
//   * by tests/ to pin the C restatement (oracle/nfl_oracle.c) and the CUDA path bit-for-bit, and
//   * by bench.py's cpu_baseline / --impl reference leg (kind = "reference").
// The product (nfllib_b200/) never loads this library.
//
// Every operation goes through the reference's own public surface:
//   fwd        -> nfl::poly::ntt_pow_phi()            include/nfl/poly.hpp:167
//   inv        -> nfl::poly::invntt_pow_invphi()      include/nfl/poly.hpp:168
//   mul/add/sub-> operator* / + / -                   include/nfl/poly.hpp:346-350
//   mul_shoup  -> nfl::shoup(a * b, bprime)           include/nfl/ops.hpp:266-277
//   compute_shoup -> nfl::compute_shoup(a)            include/nfl/poly.hpp:352
//   raw_ntt / raw_intt -> poly::core::ntt / inv_ntt via the friend proxy, as tests/ntt_perfs.cpp:122-134 does
//   polymul    -> fwd(a), fwd(b), a*b, inv            (tests/nfllib_demo_main_op.cpp:31-45 pattern)
#include <array>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <nfl.hpp>

namespace nfl { namespace tests {
// Same friend-proxy trick as tests/ntt_perfs.cpp:122-134 (poly.hpp:71-76 declares the friend).
template <class P>
class poly_tests_proxy {
public:
  static void raw_ntt(P &p, size_t cm) {
    P::core::ntt(&p(cm, 0), p.base.omegas[cm], p.base.shoupomegas[cm], P::get_modulus(cm));
  }
  static void raw_inv_ntt(P &p, size_t cm) {  // core.hpp:539-557, reached the same way
    P::core::inv_ntt(&p(cm, 0), p.base.invomegas[cm], p.base.shoupinvomegas[cm], p.base.invpolyDegree[cm], P::get_modulus(cm));
  }
};
}}

namespace {

enum Op { OP_FWD = 0, OP_INV, OP_MUL, OP_MUL_SHOUP, OP_COMPUTE_SHOUP, OP_ADD, OP_SUB, OP_RAW_NTT, OP_POLYMUL,
          OP_MULADD /* out = a + b*c, expression-fused */, OP_RAW_INTT };

template <class P>
void *aligned_polys(size_t n) {
  void *ptr = nullptr;
  if (posix_memalign(&ptr, 32, n * sizeof(P)) != 0) return nullptr;
  return ptr;
}

// One thread's share: polys [lo, hi).  Buffers are raw [batch][M][N] arrays == arrays of poly (poly.hpp:87-88).
template <class P>
void run_range(int op, P *out, const P *a, const P *b, const P *c, size_t lo, size_t hi) {
  for (size_t i = lo; i < hi; ++i) {
    switch (op) {
      case OP_FWD: out[i].ntt_pow_phi(); break;          // run_config already copied a -> out
      case OP_INV: out[i].invntt_pow_invphi(); break;
      case OP_MUL: out[i] = a[i] * b[i]; break;
      case OP_MUL_SHOUP: out[i] = nfl::shoup(a[i] * b[i], c[i]); break;
      case OP_COMPUTE_SHOUP: out[i] = nfl::compute_shoup(a[i]); break;
      case OP_ADD: out[i] = a[i] + b[i]; break;
      case OP_SUB: out[i] = a[i] - b[i]; break;
      case OP_MULADD: out[i] = a[i] + b[i] * c[i]; break;
      case OP_RAW_NTT:
        if (out != a) std::memcpy(&out[i], &a[i], sizeof(P));
        for (size_t cm = 0; cm < P::nmoduli; ++cm) nfl::tests::poly_tests_proxy<P>::raw_ntt(out[i], cm);
        break;
      case OP_RAW_INTT:
        if (out != a) std::memcpy(&out[i], &a[i], sizeof(P));
        for (size_t cm = 0; cm < P::nmoduli; ++cm) nfl::tests::poly_tests_proxy<P>::raw_inv_ntt(out[i], cm);
        break;
      case OP_POLYMUL: {
        P *ta = static_cast<P *>(aligned_polys<P>(2));
        std::memcpy(&ta[0], &a[i], sizeof(P));
        std::memcpy(&ta[1], &b[i], sizeof(P));
        ta[0].ntt_pow_phi();
        ta[1].ntt_pow_phi();
        out[i] = ta[0] * ta[1];
        out[i].invntt_pow_invphi();
        free(ta);
      } break;
    }
  }
}

template <class P>
int run_config(int op, void *out, const void *a, const void *b, const void *c, size_t batch, int threads) {
  P *po = static_cast<P *>(out);
  const P *pa = static_cast<const P *>(a), *pb = static_cast<const P *>(b), *pc = static_cast<const P *>(c);
  if ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
       reinterpret_cast<uintptr_t>(c)) & 31)
    return -2;  // reference asserts 32-byte alignment (core.hpp:88)
  if (op == OP_FWD || op == OP_INV) {
    if (out != a) std::memcpy(out, a, batch * sizeof(P));
    pa = po;
  }
  if (threads <= 1) {
    run_range<P>(op, po, pa, pb, pc, 0, batch);
    return 0;
  }
  std::vector<std::thread> pool;
  size_t per = (batch + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    size_t lo = t * per, hi = std::min(batch, lo + per);
    if (lo >= hi) break;
    pool.emplace_back(run_range<P>, op, po, pa, pb, pc, lo, hi);
  }
  for (auto &th : pool) th.join();
  return 0;
}

}  // namespace

#define NFLREF_CONFIG(T, BITS, N, M) \
  if (limb_bits == BITS && degree == N && nmoduli == M) \
    return run_config<nfl::poly<T, N, M>>(op, out, a, b, c, batch, threads);

#define NFLREF_CAT2(a, b) a##b
#define NFLREF_CAT(a, b) NFLREF_CAT2(a, b)

extern "C" {

// The instantiation list is split over several translation units (NFLREF_PART = 0..5) only so that
// `make -j` can build them in parallel; ref_dispatch.cpp tries each part in turn.
// Returns 0 on success, -1 if (limb_bits, degree, nmoduli) is not in this part, -2 on misaligned buffers.
int NFLREF_CAT(nflref_run_part, NFLREF_PART)(int op, int limb_bits, size_t degree, size_t nmoduli, void *out,
                                             const void *a, const void *b, const void *c, size_t batch,
                                             int threads) {
#include "ref_configs.inc"
  return -1;
}

#if NFLREF_PART == 2
}  // extern "C"
// CRT lift through the reference's own GMP code (gmp.hpp:183-219), linked against the image's libgmp runtime.
// words: [batch][degree][W] little-endian 64-bit words of each lifted coefficient (mpz_export, least significant first).
template <class P> static int lift_config(int dir, void *polys, uint64_t *words, size_t W, size_t batch) {
  P *p = static_cast<P *>(polys);
  std::array<mpz_t, P::degree> *arr = new std::array<mpz_t, P::degree>;
  for (size_t i = 0; i < P::degree; ++i) mpz_init((*arr)[i]);
  for (size_t b = 0; b < batch; ++b) {
    if (dir == 0) {  // poly2mpz
      p[b].poly2mpz(*arr);
      for (size_t i = 0; i < P::degree; ++i) {
        uint64_t *w = words + (b * P::degree + i) * W;
        std::memset(w, 0, W * 8);
        size_t count = 0;
        if (mpz_sizeinbase((*arr)[i], 2) > 64 * W) return -3;
        mpz_export(w, &count, -1, 8, 0, 0, (*arr)[i]);
      }
    } else {  // mpz2poly
      for (size_t i = 0; i < P::degree; ++i) mpz_import((*arr)[i], W, -1, 8, 0, 0, words + (b * P::degree + i) * W);
      p[b].mpz2poly(*arr);
    }
  }
  for (size_t i = 0; i < P::degree; ++i) mpz_clear((*arr)[i]);
  delete arr;
  return 0;
}
#define NFLREF_LIFT(T, BITS, N, M) \
  if (limb_bits == BITS && degree == N && nmoduli == M) return lift_config<nfl::poly<T, N, M>>(dir, polys, words, W, batch);
extern "C" {
// dir 0: words = poly2mpz(polys);  dir 1: polys = mpz2poly(words).  Returns -1 for a configuration that is not built.
int nflref_lift(int dir, int limb_bits, size_t degree, size_t nmoduli, void *polys, uint64_t *words, size_t W, size_t batch) {
  if (reinterpret_cast<uintptr_t>(polys) & 31) return -2;
  NFLREF_LIFT(uint64_t, 64, 1024, 4) NFLREF_LIFT(uint64_t, 64, 64, 3) NFLREF_LIFT(uint64_t, 64, 1024, 2)
  NFLREF_LIFT(uint32_t, 32, 1024, 2) NFLREF_LIFT(uint32_t, 32, 4096, 14) NFLREF_LIFT(uint16_t, 16, 512, 2)
  return -1;
}
#endif

#if NFLREF_PART == 0

}  // extern "C"
// Deterministic stand-in for the reference's entropy source lib/prng/randombytes.cpp (reads /dev/urandom): the Salsa20
// key that lib/prng/fastrandombytes.cpp:24-27 draws once becomes the fixed bytes 1, 2, ..., 32, which makes the
// reference's samplers reproducible so that poly::set(uniform) (core.hpp:150-187) can serve as an oracle.  Everything
// downstream — fastrandombytes.cpp and the Salsa20 assembly — is the reference's own, unmodified code.
namespace nfl {
void randombytes(unsigned char *x, unsigned long long xlen) {
  for (unsigned long long i = 0; i < xlen; ++i) x[i] = (unsigned char)(i + 1);
}
}
static unsigned long long g_uniform_calls = 0;  // mirrors the nonce counter inside fastrandombytes.cpp:17-34
extern "C" unsigned long long *nflref_nonce_counter(void) { return &g_uniform_calls; }  // shared with ref_gaussian.cpp
template <class P> static void sample_range(int kind, P *out, size_t batch, unsigned long long p0, unsigned long long p1) {
  for (size_t i = 0; i < batch; ++i) {
    if (kind == 0) out[i].set(nfl::uniform());                       // core.hpp:150-187
    else if (kind == 1) out[i].set(nfl::non_uniform(p0, p1));         // core.hpp:190-278
    else if (kind == 2) out[i].set(nfl::ZO_dist((uint8_t)p0));        // core.hpp:338-349
    else {                                                            // core.hpp:355-392
      out[i].set(nfl::hwt_dist((uint32_t)p0));
      // one fastrandombytes call per refill of hwt words + one for the signs (rejections have probability ~k/2^64: none)
      g_uniform_calls += (P::degree - p0 + p0 - 1) / p0;
    }
    ++g_uniform_calls;  // uniform / non_uniform / ZO: exactly one fastrandombytes call; hwt: the sign call
  }
}
#define NFLREF_UNIFORM(T, BITS, N, M) \
  if (limb_bits == BITS && degree == N && nmoduli == M) { sample_range(kind, static_cast<nfl::poly<T, N, M> *>(out), batch, p0, p1); return 0; }
extern "C" {

// out[0..batch) = successive poly::set(nfl::uniform()) draws; *first_nonce receives the 64-bit nonce the first of
// them used (one fastrandombytes call, i.e. one nonce, per polynomial).  Single-threaded: the reference's PRNG state
// is a process-global static.
// kind: 0 uniform, 1 non_uniform(p0 = upper_bound, p1 = amplifier), 2 ZO_dist(p0 = rho), 3 hwt_dist(p0 = hwt)
int nflref_sample(int kind, int limb_bits, size_t degree, size_t nmoduli, void *out, size_t batch, unsigned long long p0,
                  unsigned long long p1, unsigned long long *first_nonce) {
  if (reinterpret_cast<uintptr_t>(out) & 31) return -2;
  *first_nonce = g_uniform_calls;
  NFLREF_UNIFORM(uint64_t, 64, 1024, 4) NFLREF_UNIFORM(uint64_t, 64, 64, 3) NFLREF_UNIFORM(uint32_t, 32, 4096, 1)
  NFLREF_UNIFORM(uint32_t, 32, 8, 2) NFLREF_UNIFORM(uint16_t, 16, 512, 2) NFLREF_UNIFORM(uint16_t, 16, 16, 1)
  return -1;
}

// Tables of the reference, for pinning our own parameter derivation (include/nfl/params.hpp).
int nflref_params(int limb_bits, size_t count, uint64_t *P, uint64_t *Pn, uint64_t *roots, uint64_t *invkmax,
                  uint64_t *kmax, uint64_t *maxmoduli) {
#define DUMP(T) \
  { *kmax = nfl::params<T>::kMaxPolyDegree; *maxmoduli = nfl::params<T>::kMaxNbModuli; \
    for (size_t i = 0; i < count && i < nfl::params<T>::kMaxNbModuli; ++i) { \
      P[i] = nfl::params<T>::P[i]; Pn[i] = nfl::params<T>::Pn[i]; \
      roots[i] = nfl::params<T>::primitive_roots[i]; invkmax[i] = nfl::params<T>::invkMaxPolyDegree[i]; } \
    return 0; }
  if (limb_bits == 16) DUMP(uint16_t)
  if (limb_bits == 32) DUMP(uint32_t)
  if (limb_bits == 64) DUMP(uint64_t)
  return -1;
}

const char *nflref_build_flags(void) {
#if defined(NTT_AVX2)
  return "NFL_OPTIMIZED NTT_AVX2";
#elif defined(NTT_SSE)
  return "NFL_OPTIMIZED NTT_SSE";
#else
  return "serial";
#endif
}

#endif  // NFLREF_PART == 0

}  // extern "C"

// nfl_b200.hpp — C++11 host surface of the B200-native NFLlib hot path.
//
// Keeps the reference's template surface for the NTT / pointwise path so user code written against
// <nfl.hpp> keeps compiling for that path:
//     nfl::poly<T, Degree, NbModuli>          include/nfl/poly.hpp:82-309   (same POD layout, 32-byte aligned)
//     .ntt_pow_phi() / .invntt_pow_invphi()   poly.hpp:167-168
//     operator+ - * == !=, shoup(), compute_shoup()   poly.hpp:346-352, ops.hpp:18-45 (lazy expressions)
//     nfl::add / sub / mul, poly_from_modulus          poly.hpp:314-337
//     nfl::params<T>                                   params.hpp (tables re-derived by libnflgpu)
// but evaluates on the GPU through the C ABI of include/nflgpu.h (the only thing this header links against).
// Error convention follows the reference: std::runtime_error (core.hpp:111-115) — every nonzero nflgpu status
// is converted to one.  There is no CPU fallback: without libnflgpu + a CUDA device the calls throw.
//
// A single host-resident poly per call is PCIe-bound (hard part 5); throughput code should
// use nfl::cuda::batch<poly> below, which keeps `count` polys resident in HBM between operations.
//
// Also here: poly_p (copy-on-write handle, poly_p.hpp), serialize_manually / deserialize_manually (poly.hpp:180-191; the byte
// layout is identical, so reference-serialized polys load unchanged), the samplers behind poly::set(...), and the CRT lift
// as batch::poly2words / words2poly (GMP types themselves stay out: the lift works on little-endian 64-bit words).
// Out of scope: FastGaussianNoise::getNoise into free-standing arrays, cereal serialization.
#ifndef NFL_B200_HPP
#define NFL_B200_HPP

#include <algorithm>
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <iostream>
#include <iterator>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>

#include "nflgpu.h"

namespace nfl {

// ---------------------------------------------------------------------------------------------------------
// params<T>  (params.hpp:12-119).  The tables are fetched once from libnflgpu's derivation.
// ---------------------------------------------------------------------------------------------------------
namespace detail {

inline void check(int rc, const char *what) {
  if (rc != 0) throw std::runtime_error(std::string("nfl_b200: ") + what + ": " + nflgpu_last_error());
}

template <class T> struct limb_traits;
template <> struct limb_traits<uint16_t> {
  typedef int16_t signed_type; typedef uint32_t greater_type;
  static constexpr int bits = 16; static constexpr unsigned max_moduli = 2; static constexpr unsigned max_degree = 512;
};
template <> struct limb_traits<uint32_t> {
  typedef int32_t signed_type; typedef uint64_t greater_type;
  static constexpr int bits = 32; static constexpr unsigned max_moduli = 291; static constexpr unsigned max_degree = 32768;
};
template <> struct limb_traits<uint64_t> {
  typedef int64_t signed_type; typedef unsigned __int128 greater_type;
  static constexpr int bits = 64; static constexpr unsigned max_moduli = 1000; static constexpr unsigned max_degree = 1048576;
};

template <class T> struct param_tables {
  std::vector<uint64_t> P, Pn, roots, invkmax;
  param_tables() {
    const size_t n = limb_traits<T>::max_moduli;
    P.resize(n); Pn.resize(n); roots.resize(n); invkmax.resize(n);
    // roots of all 1000 64-bit moduli take a moment to derive; fetch P/Pn/invkmax eagerly, roots lazily
    check(nflgpu_params(limb_traits<T>::bits, 0, n, P.data(), Pn.data(), nullptr, invkmax.data()), "nflgpu_params");
    std::fill(roots.begin(), roots.end(), 0);
  }
  static param_tables &get() { static param_tables t; return t; }
  uint64_t root(size_t i) {
    if (!roots[i]) check(nflgpu_params(limb_traits<T>::bits, i, 1, nullptr, nullptr, &roots[i], nullptr), "nflgpu_params");
    return roots[i];
  }
};

template <class T, int WHICH> struct param_view {  // params<T>::P[i] syntax
  T operator[](size_t i) const {
    param_tables<T> &t = param_tables<T>::get();
    return static_cast<T>(WHICH == 0 ? t.P[i] : WHICH == 1 ? t.Pn[i] : WHICH == 2 ? t.root(i) : t.invkmax[i]);
  }
};

}  // namespace detail

template <class T> struct params {
  typedef T value_type;
  typedef typename detail::limb_traits<T>::signed_type signed_value_type;
  typedef typename detail::limb_traits<T>::greater_type greater_value_type;
  typedef value_type *poly_t;
  static constexpr unsigned int kMaxNbModuli = detail::limb_traits<T>::max_moduli;
  static constexpr unsigned int kModulusBitsize = detail::limb_traits<T>::bits - 2;
  static constexpr unsigned int kModulusRepresentationBitsize = detail::limb_traits<T>::bits;
  static constexpr unsigned int kMaxPolyDegree = detail::limb_traits<T>::max_degree;
  static const detail::param_view<T, 0> P;
  static const detail::param_view<T, 1> Pn;
  static const detail::param_view<T, 2> primitive_roots;
  static const detail::param_view<T, 3> invkMaxPolyDegree;
};
template <class T> const detail::param_view<T, 0> params<T>::P = {};
template <class T> const detail::param_view<T, 1> params<T>::Pn = {};
template <class T> const detail::param_view<T, 2> params<T>::primitive_roots = {};
template <class T> const detail::param_view<T, 3> params<T>::invkMaxPolyDegree = {};
template <class T> constexpr unsigned int params<T>::kMaxNbModuli;
template <class T> constexpr unsigned int params<T>::kModulusBitsize;
template <class T> constexpr unsigned int params<T>::kModulusRepresentationBitsize;
template <class T> constexpr unsigned int params<T>::kMaxPolyDegree;

// The reference's backend seam is the SIMD tag chosen by CC_SIMD (arch.hpp:6-18); this backend's tag:
namespace simd { struct cuda {}; struct serial {}; }  // (serial: accepted where reference code names it, tests/nfllib_demo_main_op.cpp:79)
#define CC_SIMD nfl::simd::cuda

/* Generators to initialise random polynomials (poly.hpp:42-62).  All four are drawn ON THE DEVICE by the samplers of
 * libnflgpu (nflgpu_uniform / non_uniform / zo / hwt), which follow the reference's algorithms over the same Salsa20/20
 * keystream (core.hpp:150-392, lib/prng/fastrandombytes.cpp:21-34); like the reference, the process keys itself once
 * from /dev/urandom and uses a 64-bit nonce counter. */
struct uniform {};
struct non_uniform {
  uint64_t upper_bound;
  uint64_t amplifier;
  non_uniform(uint64_t ub) : upper_bound(ub), amplifier(1) {}
  non_uniform(uint64_t ub, uint64_t amp) : upper_bound(ub), amplifier(amp) {}
};
struct hwt_dist {  // hamming weight distribution
  uint32_t hwt;
  hwt_dist(uint32_t hwt_) : hwt(hwt_) {}
};
struct ZO_dist {  // P(1) = P(-1) = (rho/0xFF)/2, P(0) = 1 - P(1) - P(-1)
  uint8_t rho;
  ZO_dist(uint8_t rho_ = 0x7F) : rho(rho_) {}
};

/* Discrete Gaussian (poly.hpp:61-67, prng/FastGaussianNoise.hpp).  The object holds what the reference's constructor builds
 * (barrier table in MPFR arithmetic, look-up tables), built by libnflgpu on first use on the device of the polynomial type that
 * draws from it; poly::set(gaussian) / batch::set_gaussian run the look-up sampler on the device (nflgpu_gaussian_sample),
 * bit-identical to the reference's draws under the same key and nonce.  Shapes: (uint8_t, 1 | 2), (uint16_t, 1).
 * Not mirrored: getNoise() into a free-standing array (only polynomials are drawn), verbose output. */
template <class in_class, class out_class, unsigned _lu_depth> class FastGaussianNoise {
  static_assert((sizeof(in_class) == 1 && (_lu_depth == 1 || _lu_depth == 2)) || (sizeof(in_class) == 2 && _lu_depth == 1),
                "FastGaussianNoise: in_class must be uint8_t (depth 1 or 2) or uint16_t (depth 1)");
  double sigma_, center_;
  unsigned security_, samples_;
  nflgpu_gaussian *g_;
  FastGaussianNoise(FastGaussianNoise const &);
  FastGaussianNoise &operator=(FastGaussianNoise const &);

public:
  FastGaussianNoise(double sigma, unsigned int security, unsigned int samples, double center_d = 0, bool /*verbose*/ = false)
      : sigma_(sigma), center_(center_d), security_(security), samples_(samples), g_(nullptr) {}
  ~FastGaussianNoise() { if (g_) nflgpu_gaussian_destroy(g_); }
  nflgpu_gaussian *handle(nflgpu_ctx *ctx) {
    if (!g_ && nflgpu_gaussian_create(&g_, ctx, sigma_, security_, samples_, center_, (int)sizeof(in_class), (int)_lu_depth) != NFLGPU_OK)
      throw std::runtime_error(std::string("FastGaussianNoise: ") + nflgpu_last_error());
    return g_;
  }
};
template <class in_class, class out_class, unsigned _lu_depth> struct gaussian {
  FastGaussianNoise<in_class, out_class, _lu_depth> *fg_prng;
  uint64_t amplifier;
  gaussian(FastGaussianNoise<in_class, out_class, _lu_depth> *prng) : fg_prng(prng), amplifier(1) {}
  gaussian(FastGaussianNoise<in_class, out_class, _lu_depth> *prng, uint64_t amp) : fg_prng(prng), amplifier(amp) {}
};

namespace detail {
struct prng_state {  // lib/prng/fastrandombytes.cpp:17-34: key from the OS once, nonce counter from 0
  uint8_t key[32];
  std::atomic<uint64_t> nonce;
  prng_state() : nonce(0) {
    FILE *f = std::fopen("/dev/urandom", "rb");
    const bool ok = f && std::fread(key, 1, sizeof(key), f) == sizeof(key);
    if (f) std::fclose(f);
    if (!ok) throw std::runtime_error("nfl_b200: cannot read /dev/urandom");
  }
  static prng_state &get() { static prng_state s; return s; }
  uint64_t take(uint64_t count) { return nonce.fetch_add(count); }
};
}  // namespace detail

template <class T, size_t Degree, size_t NbModuli> class poly;
template <class T, size_t Degree, size_t NbModuli> class poly_p;

// proxy the reference's tests use to reach poly's protected members (poly.hpp:71-76,85,202; tests/ntt_perfs.cpp:122-134)
namespace tests { template <class P> class poly_tests_proxy; }

// ---------------------------------------------------------------------------------------------------------
// Backend: one nflgpu context per (T, Degree, NbModuli), created on first use (the reference builds its
// tables at static-init time, core.hpp:45-62).  Device ordinal: environment variable NFL_B200_DEVICE (default 0).
// ---------------------------------------------------------------------------------------------------------
namespace detail {

template <class T, size_t Degree, size_t NbModuli> struct backend {
  nflgpu_ctx *ctx;
  backend() : ctx(nullptr) {
    static_assert(Degree <= limb_traits<T>::max_degree, "degree > params<T>::kMaxPolyDegree");    // core.hpp:59-60
    static_assert(NbModuli <= limb_traits<T>::max_moduli, "nmoduli > params<T>::kMaxNbModuli");   // core.hpp:57-58
    static_assert((Degree & (Degree - 1)) == 0 && Degree * sizeof(T) >= 32, "degree must be a power of two, >= 32 bytes of limbs");
    const char *dev = std::getenv("NFL_B200_DEVICE");
    check(nflgpu_ctx_create(&ctx, limb_traits<T>::bits, Degree, NbModuli, 0, dev ? std::atoi(dev) : 0, nullptr, nullptr),
          "nflgpu_ctx_create");
  }
  ~backend() { nflgpu_ctx_destroy(ctx); }
  static backend &get() { static backend b; return b; }
};

// One single-residue context per modulus index, for the per-residue statics core::ntt / core::inv_ntt (core.hpp:455-557)
template <class T, size_t Degree> struct residue_backend {
  std::mutex mu;
  std::map<size_t, nflgpu_ctx *> ctxs;
  ~residue_backend() { for (auto &kv : ctxs) nflgpu_ctx_destroy(kv.second); }
  static residue_backend &get() { static residue_backend b; return b; }
  nflgpu_ctx *ctx(size_t cm) {
    std::lock_guard<std::mutex> lock(mu);
    auto it = ctxs.find(cm);
    if (it != ctxs.end()) return it->second;
    const char *dev = std::getenv("NFL_B200_DEVICE");
    nflgpu_ctx *c = nullptr;
    check(nflgpu_ctx_create(&c, limb_traits<T>::bits, Degree, 1, cm, dev ? std::atoi(dev) : 0, nullptr, nullptr), "nflgpu_ctx_create");
    ctxs[cm] = c;
    return c;
  }
};

// RAII device buffer of `count` polys.  Temporaries of single-poly calls (pooled = true, the default) come from the
// context's stream-ordered pool (nflgpu_scratch_alloc: a cached block, no cudaMalloc / cudaFree per leaf of an expression);
// long-lived batches (nfl::cuda::batch) own a plain allocation that can also be exported to a peer process.
template <class P> struct dev_buf {
  void *p; size_t count; bool pooled;
  explicit dev_buf(size_t n, bool pooled_ = true) : p(nullptr), count(n), pooled(pooled_) {
    if (pooled) check(nflgpu_scratch_alloc(P::backend_type::get().ctx, n, &p, nullptr), "nflgpu_scratch_alloc");
    else check(nflgpu_alloc(P::backend_type::get().ctx, n, &p), "nflgpu_alloc");
  }
  ~dev_buf() {
    if (!p) return;
    if (pooled) nflgpu_scratch_free(P::backend_type::get().ctx, p, nullptr);
    else nflgpu_free(P::backend_type::get().ctx, p);
  }
  dev_buf(const dev_buf &) = delete;
  dev_buf &operator=(const dev_buf &) = delete;
};

}  // namespace detail

// ---------------------------------------------------------------------------------------------------------
// Expression templates (ops.hpp:47-277).  A whole tree is flattened into a postfix program and evaluated by ONE kernel
// (nflgpu_eval): every leaf is uploaded and read once, as in the reference's fused evaluator loop; `shoup(a*b, b')` is
// rewritten to mulmod_shoup (ops.hpp:266-277).  Trees beyond nflgpu_eval's limits fall back to one kernel per node.
// ---------------------------------------------------------------------------------------------------------
namespace ops {

// (class templates with the reference's parameter list <value type, SIMD tag> — ops.hpp:124-242 — so that reference code naming
//  a functor explicitly, e.g. ops::make_op<ops::mulmod_shoup<T, simd::serial>>(a, b, b') in tests/nfllib_demo_main_op.cpp:79,
//  compiles; the parameters select nothing here: every functor runs on the device)
template <class T = void, class Tag = void> struct addmod {};
template <class T = void, class Tag = void> struct submod {};
template <class T = void, class Tag = void> struct mulmod {};
template <class T = void, class Tag = void> struct mulmod_shoup {};
template <class T = void, class Tag = void> struct compute_shoup {};
struct shoup {};
struct eqmod {}; struct neqmod {};

template <class Op, class... Args> struct expr {
  typedef typename std::tuple_element<0, std::tuple<Args...>>::type first_arg;
  typedef typename first_arg::poly_type poly_type;
  std::tuple<Args const &...> args;
  explicit expr(Args const &... a) : args(a...) {}
  // the reference's semantics: `==` is true iff ANY coefficient is equal, `!=` iff ANY differs (ops.hpp:81-95);
  // implicit like the reference's, so `ret &= (a == b)` and `bool ok = (a == b)` compile (tests/poly_p.cpp:22)
  operator bool() const;
};

}  // namespace ops

namespace detail {

template <class P> struct eval;  // forward

// evaluates any operand (poly or expr) into a fresh device buffer of one poly
template <class P, class X> struct operand_eval;
template <class P> struct operand_eval<P, P> {
  static std::unique_ptr<dev_buf<P>> run(P const &x) {
    std::unique_ptr<dev_buf<P>> b(new dev_buf<P>(1));
    check(nflgpu_upload(P::backend_type::get().ctx, b->p, x.data(), 1, nullptr), "nflgpu_upload");
    return b;
  }
};
template <class T, size_t D, size_t M> struct operand_eval<poly<T, D, M>, poly_p<T, D, M>> {
  static std::unique_ptr<dev_buf<poly<T, D, M>>> run(poly_p<T, D, M> const &x) { return operand_eval<poly<T, D, M>, poly<T, D, M>>::run(x.poly_obj()); }
};

#define NFLB200_BIN(OPTAG, CALL)                                                                                    \
  template <class P, class A0, class A1> struct operand_eval<P, ops::expr<ops::OPTAG, A0, A1>> {                   \
    static std::unique_ptr<dev_buf<P>> run(ops::expr<ops::OPTAG, A0, A1> const &e) {                               \
      auto a = operand_eval<P, A0>::run(std::get<0>(e.args));                                                       \
      auto b = operand_eval<P, A1>::run(std::get<1>(e.args));                                                       \
      check(CALL(P::backend_type::get().ctx, a->p, a->p, b->p, 1, nullptr), #CALL);                                \
      return a;                                                                                                     \
    }                                                                                                               \
  };
NFLB200_BIN(submod<>, nflgpu_sub)
NFLB200_BIN(mulmod<>, nflgpu_mul)
#undef NFLB200_BIN

// x + y  — with the fused form when y is a product (core.hpp:24-37 evaluates the whole tree in one pass)
template <class P, class A0, class A1> struct operand_eval<P, ops::expr<ops::addmod<>, A0, A1>> {
  static std::unique_ptr<dev_buf<P>> run(ops::expr<ops::addmod<>, A0, A1> const &e) {
    auto a = operand_eval<P, A0>::run(std::get<0>(e.args));
    auto b = operand_eval<P, A1>::run(std::get<1>(e.args));
    check(nflgpu_add(P::backend_type::get().ctx, a->p, a->p, b->p, 1, nullptr), "nflgpu_add");
    return a;
  }
};
template <class P, class A0, class B0, class B1> struct operand_eval<P, ops::expr<ops::addmod<>, A0, ops::expr<ops::mulmod<>, B0, B1>>> {
  static std::unique_ptr<dev_buf<P>> run(ops::expr<ops::addmod<>, A0, ops::expr<ops::mulmod<>, B0, B1>> const &e) {
    auto a = operand_eval<P, A0>::run(std::get<0>(e.args));
    auto const &m = std::get<1>(e.args);
    auto b = operand_eval<P, B0>::run(std::get<0>(m.args));
    auto c = operand_eval<P, B1>::run(std::get<1>(m.args));
    check(nflgpu_muladd(P::backend_type::get().ctx, a->p, a->p, b->p, c->p, 1, nullptr), "nflgpu_muladd");
    return a;
  }
};
template <class P, class A0, class A1, class A2> struct operand_eval<P, ops::expr<ops::mulmod_shoup<>, A0, A1, A2>> {
  static std::unique_ptr<dev_buf<P>> run(ops::expr<ops::mulmod_shoup<>, A0, A1, A2> const &e) {
    auto a = operand_eval<P, A0>::run(std::get<0>(e.args));
    auto b = operand_eval<P, A1>::run(std::get<1>(e.args));
    auto c = operand_eval<P, A2>::run(std::get<2>(e.args));
    check(nflgpu_mul_shoup(P::backend_type::get().ctx, a->p, a->p, b->p, c->p, 1, nullptr), "nflgpu_mul_shoup");
    return a;
  }
};
template <class P, class A0> struct operand_eval<P, ops::expr<ops::compute_shoup<>, A0>> {
  static std::unique_ptr<dev_buf<P>> run(ops::expr<ops::compute_shoup<>, A0> const &e) {
    auto a = operand_eval<P, A0>::run(std::get<0>(e.args));
    check(nflgpu_compute_shoup(P::backend_type::get().ctx, a->p, a->p, 1, nullptr), "nflgpu_compute_shoup");
    return a;
  }
};

// ---- fused path: flatten the tree into the postfix program of nflgpu_eval (one kernel, every leaf read once) ----
template <class P> struct rpn {
  std::vector<const P *> leaves;
  std::vector<uint8_t> prog;
  int depth, max_depth;
  bool ok;
  rpn() : depth(0), max_depth(0), ok(true) {}
  void leaf(const P &p) {
    size_t i = 0;
    while (i < leaves.size() && leaves[i] != &p) ++i;   // the same poly appearing twice is uploaded once
    if (i == leaves.size()) { if (leaves.size() == 8) { ok = false; return; } leaves.push_back(&p); }
    prog.push_back(static_cast<uint8_t>(i));
    if (++depth > max_depth) max_depth = depth;
  }
  void op(uint8_t tok, int pops) { prog.push_back(tok); depth -= pops - 1; }
};
template <class P, class X> struct emit;
template <class P> struct emit<P, P> { static void run(rpn<P> &r, P const &x) { r.leaf(x); } };
template <class T, size_t D, size_t M> struct emit<poly<T, D, M>, poly_p<T, D, M>> {
  static void run(rpn<poly<T, D, M>> &r, poly_p<T, D, M> const &x) { r.leaf(x.poly_obj()); }
};
#define NFLB200_EMIT_BIN(OPTAG, TOK)                                                             \
  template <class P, class A0, class A1> struct emit<P, ops::expr<ops::OPTAG, A0, A1>> {        \
    static void run(rpn<P> &r, ops::expr<ops::OPTAG, A0, A1> const &e) {                         \
      emit<P, A0>::run(r, std::get<0>(e.args)); emit<P, A1>::run(r, std::get<1>(e.args)); r.op(TOK, 2); \
    }                                                                                            \
  };
NFLB200_EMIT_BIN(addmod<>, 0x10)
NFLB200_EMIT_BIN(submod<>, 0x11)
NFLB200_EMIT_BIN(mulmod<>, 0x12)
#undef NFLB200_EMIT_BIN
template <class P, class A0, class A1, class A2> struct emit<P, ops::expr<ops::mulmod_shoup<>, A0, A1, A2>> {
  static void run(rpn<P> &r, ops::expr<ops::mulmod_shoup<>, A0, A1, A2> const &e) {
    emit<P, A0>::run(r, std::get<0>(e.args)); emit<P, A1>::run(r, std::get<1>(e.args)); emit<P, A2>::run(r, std::get<2>(e.args));
    r.op(0x13, 3);
  }
};
template <class P, class A0> struct emit<P, ops::expr<ops::compute_shoup<>, A0>> {
  static void run(rpn<P> &r, ops::expr<ops::compute_shoup<>, A0> const &e) { emit<P, A0>::run(r, std::get<0>(e.args)); r.op(0x14, 1); }
};

// Evaluates `e` into host poly `out`: fused single kernel when the tree fits nflgpu_eval's limits, node by node otherwise.
template <class P, class X> void assign_expr(P &out, X const &e) {
  nflgpu_ctx *ctx = P::backend_type::get().ctx;
  rpn<P> r;
  emit<P, X>::run(r, e);
  if (r.ok && r.prog.size() <= 32 && r.max_depth <= 8) {
    std::vector<std::unique_ptr<dev_buf<P>>> bufs;
    std::vector<const void *> ptrs;
    for (size_t i = 0; i < r.leaves.size(); ++i) {
      bufs.emplace_back(new dev_buf<P>(1));
      check(nflgpu_upload(ctx, bufs.back()->p, r.leaves[i]->data(), 1, nullptr), "nflgpu_upload");
      ptrs.push_back(bufs.back()->p);
    }
    check(nflgpu_eval(ctx, bufs[0]->p, ptrs.data(), ptrs.size(), r.prog.data(), r.prog.size(), 1, nullptr), "nflgpu_eval");
    check(nflgpu_download(ctx, out.data(), bufs[0]->p, 1, nullptr), "nflgpu_download");
    check(nflgpu_sync(ctx, nullptr), "nflgpu_sync");
    return;
  }
  auto res = operand_eval<P, X>::run(e);
  check(nflgpu_download(ctx, out.data(), res->p, 1, nullptr), "nflgpu_download");
  check(nflgpu_sync(ctx, nullptr), "nflgpu_sync");
}

template <class X> struct is_operand : std::false_type {};
template <class T, size_t D, size_t M> struct is_operand<poly<T, D, M>> : std::true_type {};
template <class T, size_t D, size_t M> struct is_operand<poly_p<T, D, M>> : std::true_type {};
template <class Op, class... A> struct is_operand<ops::expr<Op, A...>> : std::true_type {};

}  // namespace detail

// ---------------------------------------------------------------------------------------------------------
// poly  (poly.hpp:82-309)
// ---------------------------------------------------------------------------------------------------------
template <class T, size_t Degree, size_t NbModuli> class poly {
  template <class P> friend class tests::poly_tests_proxy;

  static constexpr size_t N = Degree * NbModuli;
  T _data[N] __attribute__((aligned(32)));

public:
  typedef poly poly_type;
  typedef detail::backend<T, Degree, NbModuli> backend_type;
  using value_type = typename params<T>::value_type;
  using greater_value_type = typename params<T>::greater_value_type;
  using signed_value_type = typename params<T>::signed_value_type;
  using pointer_type = T *;
  using const_pointer_type = T const *;
  using iterator = pointer_type;
  using const_iterator = const_pointer_type;
  using simd_mode = CC_SIMD;
  static constexpr size_t degree = Degree;
  static constexpr size_t nmoduli = NbModuli;
  static constexpr size_t nbits = params<T>::kModulusBitsize;
  static constexpr size_t aggregated_modulus_bit_size = NbModuli * nbits;

  /* constructors (core.hpp:64-147) */
  poly() { set(value_type(0)); }
  poly(uniform const &mode) { set(mode); }
  poly(non_uniform const &mode) { set(mode); }
  poly(hwt_dist const &mode) { set(mode); }
  poly(ZO_dist const &mode) { set(mode); }
  template <class in_class, unsigned _lu_depth> poly(gaussian<in_class, T, _lu_depth> const &mode) { set(mode); }
  poly(value_type v, bool reduce_coeffs = true) { set(v, reduce_coeffs); }
  poly(std::initializer_list<value_type> values, bool reduce_coeffs = true) { set(values.begin(), values.end(), reduce_coeffs); }
  template <class It> poly(It first, It last, bool reduce_coeffs = true) { set(first, last, reduce_coeffs); }
  template <class Op, class... Args> poly(ops::expr<Op, Args...> const &e) { *this = e; }

  /* set() (core.hpp:76-147): coefficient v (resp. the list) replicated over the residues, reduced mod p_cm */
  void set(value_type v, bool reduce_coeffs = true) {
    if (v == 0) { std::memset(_data, 0, sizeof(_data)); return; }
    for (size_t cm = 0; cm < nmoduli; ++cm) {
      _data[cm * degree] = reduce_coeffs ? static_cast<T>(v % get_modulus(cm)) : v;
      std::fill(_data + cm * degree + 1, _data + (cm + 1) * degree, T(0));
    }
  }
  void set(std::initializer_list<value_type> values, bool reduce_coeffs = true) { set(values.begin(), values.end(), reduce_coeffs); }
  template <class It> void set(It first, It last, bool reduce_coeffs = true) {
    // core.hpp:100-147: either `degree` values (replicated into every residue) or degree*nmoduli values
    const size_t size = std::distance(first, last);
    if (size > degree && size != degree * nmoduli)
      throw std::runtime_error("poly: CRITICAL, initializer of size above degree but not equal to nmoduli*degree");  // core.hpp:111-115
    for (size_t cm = 0; cm < nmoduli; ++cm) {
      It it = first;
      if (size == degree * nmoduli) std::advance(it, cm * degree);
      size_t i = 0;
      for (; i < degree && it != last && i < size; ++i, ++it)
        _data[cm * degree + i] = reduce_coeffs ? static_cast<T>(static_cast<value_type>(*it) % get_modulus(cm)) : static_cast<T>(*it);
      for (; i < degree; ++i) _data[cm * degree + i] = 0;
    }
  }
  /* random fills (core.hpp:150-392): one device sampler launch + download per call */
  void set(uniform const &) {
    detail::prng_state &g = detail::prng_state::get();
    detail::dev_buf<poly> b(1);
    detail::check(nflgpu_uniform(backend_type::get().ctx, b.p, 1, g.key, g.take(1), nullptr), "nflgpu_uniform");
    fetch(b);
  }
  void set(non_uniform const &mode) {
    if (mode.upper_bound >= get_modulus(0))  // core.hpp:201-206
      throw std::runtime_error("set(non_uniform): upper_bound is larger than the modulus");
    detail::prng_state &g = detail::prng_state::get();
    detail::dev_buf<poly> b(1);
    detail::check(nflgpu_non_uniform(backend_type::get().ctx, b.p, 1, mode.upper_bound, mode.amplifier, g.key, g.take(1), nullptr),
                  "nflgpu_non_uniform");
    fetch(b);
  }
  void set(hwt_dist const &mode) {
    detail::prng_state &g = detail::prng_state::get();
    detail::dev_buf<poly> b(1);
    uint64_t used = 0;  // nonces the draw consumed: data dependent in the (rare) case of a rejected index, like the reference
    detail::check(nflgpu_hwt_count(backend_type::get().ctx, b.p, 1, mode.hwt, g.key, g.nonce.load(), &used, nullptr), "nflgpu_hwt");
    g.take(used);
    fetch(b);
  }
  void set(ZO_dist const &mode) {
    detail::prng_state &g = detail::prng_state::get();
    detail::dev_buf<poly> b(1);
    detail::check(nflgpu_zo(backend_type::get().ctx, b.p, 1, mode.rho, g.key, g.take(1), nullptr), "nflgpu_zo");
    fetch(b);
  }

  /* core.hpp:291-325: a draw consumes a data-dependent number of nonces (getNoise refills); like the reference's PRNG state,
   * not thread safe */
  template <class in_class, unsigned _lu_depth> void set(gaussian<in_class, T, _lu_depth> const &mode) {
    detail::prng_state &g = detail::prng_state::get();
    detail::dev_buf<poly> b(1);
    uint64_t used = 0;
    nflgpu_ctx *ctx = backend_type::get().ctx;
    detail::check(nflgpu_gaussian_sample(ctx, mode.fg_prng->handle(ctx), b.p, 1, mode.amplifier, g.key, g.nonce.load(), &used, nullptr),
                  "nflgpu_gaussian_sample");
    g.take(used);
    fetch(b);
  }

  /* assignment */
  template <class in_class, unsigned _lu_depth> poly &operator=(gaussian<in_class, T, _lu_depth> const &mode) { set(mode); return *this; }
  poly &operator=(value_type v) { set(v); return *this; }
  poly &operator=(uniform const &mode) { set(mode); return *this; }
  poly &operator=(non_uniform const &mode) { set(mode); return *this; }
  poly &operator=(hwt_dist const &mode) { set(mode); return *this; }
  poly &operator=(ZO_dist const &mode) { set(mode); return *this; }
  poly &operator=(std::initializer_list<value_type> values) { set(values); return *this; }
  template <class Op, class... Args> poly &operator=(ops::expr<Op, Args...> const &e) {  // core.hpp:24-37
    detail::assign_expr<poly, ops::expr<Op, Args...>>(*this, e);
    return *this;
  }

  /* conversion (poly.hpp:138, core.hpp:39-43): true iff some coefficient is nonzero */
  explicit operator bool() const { return std::find_if(begin(), end(), [](value_type v) { return v != 0; }) != end(); }

  /* iterators, indexing, misc (poly.hpp:141-163) */
  iterator begin() { return _data; }
  iterator end() { return _data + N; }
  const_iterator begin() const { return _data; }
  const_iterator end() const { return _data + N; }
  const_iterator cbegin() const { return _data; }
  const_iterator cend() const { return _data + N; }
  value_type const &operator()(size_t cm, size_t i) const { return _data[cm * degree + i]; }
  value_type &operator()(size_t cm, size_t i) { return _data[cm * degree + i]; }
  pointer_type data() { return _data; }
  const_pointer_type data() const { return _data; }
  static value_type get_modulus(size_t n) { return params<T>::P[n]; }

  /* NTT — public API (poly.hpp:167-168); one host poly per call: H2D + kernel + D2H inside libnflgpu */
  void ntt_pow_phi() { detail::check(nflgpu_host_op(backend_type::get().ctx, 0, _data, _data, nullptr, nullptr, 1), "ntt_pow_phi"); }
  void invntt_pow_invphi() { detail::check(nflgpu_host_op(backend_type::get().ctx, 1, _data, _data, nullptr, nullptr, 1), "invntt_pow_invphi"); }

  /* manual serializers (poly.hpp:180-185): identical byte layout */
  void serialize_manually(std::ostream &os) { os.write(reinterpret_cast<char *>(_data), N * sizeof(T)); }
  void deserialize_manually(std::istream &is) { is.read(reinterpret_cast<char *>(_data), N * sizeof(T)); }
  /* serializer (cereal, poly.hpp:187-191) */
  template <class Archive> void serialize(Archive &archive) { archive(_data); }

private:
  void fetch(detail::dev_buf<poly> const &b) {
    detail::check(nflgpu_download(backend_type::get().ctx, _data, b.p, 1, nullptr), "nflgpu_download");
    detail::check(nflgpu_sync(backend_type::get().ctx, nullptr), "nflgpu_sync");
  }

protected:
  /* poly::core (poly.hpp:195-238): the per-residue statics and the reference's host-visible tables, kept so code that
   * reaches them through tests::poly_tests_proxy (tests/ntt_perfs.cpp:122-134) compiles unchanged.  The device kernels
   * use their own merged-psi tables (csrc/tables.cpp); these copies follow core::initialize / prep_wtab
   * (core.hpp:564-581,625-686) value for value and exist only when `base` is referenced. */
  class core {
    template <class P> friend class tests::poly_tests_proxy;

  public:
    core() { initialize(); }
    void ntt_pow_phi(poly &op) { op.ntt_pow_phi(); }
    void invntt_pow_invphi(poly &op) { op.invntt_pow_invphi(); }
    // core.hpp:455-532 — cyclic transform of ONE residue, in place, canonical in and out; always returns true
    // (core.hpp:531).  Runs on the device (nflgpu_ntt_raw_fwd); `wtab` must be this class's own table of the
    // residue with modulus p (the device tables are derived from the same omega), anything else throws.
    static bool ntt(value_type *x, const value_type *wtab, const value_type *winvtab, value_type const p) {
      if (degree == 1) return true;
      const size_t cm = residue_of(p);
      if (wtab != base.omegas[cm] || winvtab != base.shoupomegas[cm])
        throw std::runtime_error("nfl_b200: core::ntt runs with the library's own omega tables only");
      detail::check(nflgpu_host_op(detail::residue_backend<T, Degree>::get().ctx(cm), 10, x, x, nullptr, nullptr, 1), "core::ntt");
      return true;
    }
    // core.hpp:539-557 — bit-reverse, core::ntt with omega^-1, bit-reverse; invK is unused there as well
    static bool inv_ntt(value_type *x, const value_type *inv_wtab, const value_type *inv_winvtab, value_type, value_type const p) {
      if (degree == 1) return true;
      const size_t cm = residue_of(p);
      if (inv_wtab != base.invomegas[cm] || inv_winvtab != base.shoupinvomegas[cm])
        throw std::runtime_error("nfl_b200: core::inv_ntt runs with the library's own omega^-1 tables only");
      detail::check(nflgpu_host_op(detail::residue_backend<T, Degree>::get().ctx(cm), 11, x, x, nullptr, nullptr, 1), "core::inv_ntt");
      return true;
    }

  private:
    value_type phis[nmoduli][degree] __attribute__((aligned(32))), shoupphis[nmoduli][degree] __attribute__((aligned(32))),
        invpoly_times_invphis[nmoduli][degree] __attribute__((aligned(32))),
        shoupinvpoly_times_invphis[nmoduli][degree] __attribute__((aligned(32))), omegas[nmoduli][degree * 2] __attribute__((aligned(32))),
        *shoupomegas[nmoduli], invomegas[nmoduli][2 * degree] __attribute__((aligned(32))), *shoupinvomegas[nmoduli], invpolyDegree[nmoduli];

    static size_t residue_of(value_type p) {
      for (size_t cm = 0; cm < nmoduli; ++cm)
        if (get_modulus(cm) == p) return cm;
      throw std::runtime_error("nfl_b200: modulus is not one of this poly type's moduli");
    }
    static value_type mulm(value_type a, value_type b, value_type p) { return static_cast<value_type>((static_cast<unsigned __int128>(a) * b) % p); }
    static value_type shoupv(value_type v, value_type p) {  // floor(v * 2^w / p), core.hpp:575
      return static_cast<value_type>((static_cast<unsigned __int128>(v) << params<T>::kModulusRepresentationBitsize) / p);
    }
    static void prep_wtab(value_type *wtab, value_type *wtabshoup, value_type w, value_type p) {  // core.hpp:564-581
      for (size_t K = degree; K >= 2; K /= 2) {
        value_type wi = 1;
        for (size_t i = 0; i < K / 2; ++i) { *wtab++ = wi; *wtabshoup++ = shoupv(wi, p); wi = mulm(wi, w, p); }
        w = mulm(w, w, p);
      }
    }
    void initialize() {  // core.hpp:625-686
      for (size_t cm = 0; cm < nmoduli; ++cm) {
        const value_type p = get_modulus(cm);
        shoupomegas[cm] = omegas[cm] + degree;
        shoupinvomegas[cm] = invomegas[cm] + degree;
        value_type phi = params<T>::primitive_roots[cm];
        for (size_t d = degree; d < params<T>::kMaxPolyDegree; d *= 2) phi = mulm(phi, phi, p);
        value_type t = 1;
        for (size_t i = 0; i < degree; ++i) { phis[cm][i] = t; shoupphis[cm][i] = shoupv(t, p); t = mulm(t, phi, p); }
        const value_type invphi = mulm(t, phis[cm][degree - 1], p);  // phi^degree * phi^(degree-1) = phi^-1
        invpolyDegree[cm] = mulm(params<T>::invkMaxPolyDegree[cm], static_cast<value_type>(params<T>::kMaxPolyDegree / degree), p);
        t = invpolyDegree[cm];
        for (size_t i = 0; i < degree; ++i) {
          invpoly_times_invphis[cm][i] = t; shoupinvpoly_times_invphis[cm][i] = shoupv(t, p); t = mulm(t, invphi, p);
        }
        prep_wtab(omegas[cm], shoupomegas[cm], mulm(phi, phi, p), p);
        prep_wtab(invomegas[cm], shoupinvomegas[cm], mulm(invphi, invphi, p), p);
      }
    }
  } __attribute__((aligned(32)));

  static core base;
} __attribute__((aligned(32)));

template <class T, size_t D, size_t M> typename poly<T, D, M>::core poly<T, D, M>::base;

template <class T, size_t D, size_t M> constexpr size_t poly<T, D, M>::degree;
template <class T, size_t D, size_t M> constexpr size_t poly<T, D, M>::nmoduli;
template <class T, size_t D, size_t M> constexpr size_t poly<T, D, M>::nbits;
template <class T, size_t D, size_t M> constexpr size_t poly<T, D, M>::aggregated_modulus_bit_size;

// ---------------------------------------------------------------------------------------------------------
// poly_p  (poly_p.hpp:11-204): shared, copy-on-write handle to a 32-byte aligned heap poly.  Copies are O(1); the first
// mutating access through a shared handle clones the coefficients (poly_p.hpp:176-183).  Takes part in expressions
// exactly like a poly (tests/poly_p.cpp).
// ---------------------------------------------------------------------------------------------------------
template <class T, size_t Degree, size_t NbModuli> class poly_p {
public:
  typedef poly<T, Degree, NbModuli> poly_type;
  typedef typename poly_type::backend_type backend_type;
  using value_type = typename poly_type::value_type;
  using greater_value_type = typename poly_type::greater_value_type;
  static constexpr size_t nmoduli = poly_type::nmoduli;
  static constexpr size_t degree = poly_type::degree;
  static constexpr size_t nbits = poly_type::nbits;
  static constexpr size_t aggregated_modulus_bit_size = poly_type::aggregated_modulus_bit_size;

private:
  std::shared_ptr<poly_type> p_;
  template <class... Args> static std::shared_ptr<poly_type> make(Args &&... args) {
    void *raw = nullptr;
    if (posix_memalign(&raw, 32, sizeof(poly_type)) != 0) throw std::bad_alloc();
    poly_type *obj;
    try { obj = new (raw) poly_type(std::forward<Args>(args)...); } catch (...) { free(raw); throw; }
    return std::shared_ptr<poly_type>(obj, [](poly_type *q) { q->~poly_type(); free(q); });
  }
  void detach() { if (p_.use_count() > 1) p_ = make(p_->begin(), p_->end(), false); }

public:
  poly_p() : p_(make()) {}
  poly_p(poly_p const &o) : p_(o.p_) {}
  poly_p(poly_p &o) : p_(o.p_) {}
  poly_p(poly_p &&o) : p_(std::move(o.p_)) {}
  template <class A0, class... Args> poly_p(A0 &&a0, Args &&... args) : p_(make(std::forward<A0>(a0), std::forward<Args>(args)...)) {}

  poly_type &poly_obj() { detach(); return *p_; }
  poly_type const &poly_obj() const { return *p_; }

  poly_p &operator=(poly_p const &o) { p_ = o.p_; return *this; }
  poly_p &operator=(poly_p &&o) { p_ = std::move(o.p_); return *this; }
  poly_p &operator=(std::initializer_list<value_type> values) { poly_obj() = values; return *this; }
  template <class O> poly_p &operator=(O &&o) { poly_obj() = std::forward<O>(o); return *this; }

  value_type &operator()(size_t cm, size_t i) { return poly_obj()(cm, i); }
  value_type const &operator()(size_t cm, size_t i) const { return poly_obj()(cm, i); }
  typename poly_type::iterator begin() { return poly_obj().begin(); }
  typename poly_type::iterator end() { return poly_obj().end(); }
  typename poly_type::const_iterator begin() const { return poly_obj().begin(); }
  typename poly_type::const_iterator end() const { return poly_obj().end(); }
  typename poly_type::pointer_type data() { return poly_obj().data(); }
  static value_type get_modulus(size_t n) { return poly_type::get_modulus(n); }
  void ntt_pow_phi() { poly_obj().ntt_pow_phi(); }
  void invntt_pow_invphi() { poly_obj().invntt_pow_invphi(); }
  template <class... Args> void set(Args &&... args) { poly_obj().set(std::forward<Args>(args)...); }
  void serialize_manually(std::ostream &os) { poly_obj().serialize_manually(os); }
  void deserialize_manually(std::istream &is) { poly_obj().deserialize_manually(is); }
  template <class Archive> void serialize(Archive &archive) { archive(poly_obj()); }
};
template <class T, size_t D, size_t M> constexpr size_t poly_p<T, D, M>::degree;
template <class T, size_t D, size_t M> constexpr size_t poly_p<T, D, M>::nmoduli;
template <class T, size_t Degree, size_t AggregatedModulusBitSize>
using poly_p_from_modulus = poly_p<T, Degree, AggregatedModulusBitSize / params<T>::kModulusBitsize>;

/* operator overloads (poly.hpp:346-352 via the macros of ops.hpp:18-45) */
#define NFLB200_DECLARE_BINARY(NAME, TAG)                                                                                   \
  template <class A0, class A1>                                                                                             \
  typename std::enable_if<detail::is_operand<A0>::value && detail::is_operand<A1>::value, ops::expr<ops::TAG, A0, A1>>::type \
  NAME(A0 const &a, A1 const &b) { return ops::expr<ops::TAG, A0, A1>(a, b); }
NFLB200_DECLARE_BINARY(operator-, submod<>)
NFLB200_DECLARE_BINARY(operator+, addmod<>)
NFLB200_DECLARE_BINARY(operator*, mulmod<>)
NFLB200_DECLARE_BINARY(operator==, eqmod)
NFLB200_DECLARE_BINARY(operator!=, neqmod)
#undef NFLB200_DECLARE_BINARY

// ops::make_op<Functor<T, Tag>>(args...) (ops.hpp:249-262): an expression node for an explicitly named functor
namespace ops {
template <class Op> struct node_tag;
template <class T, class Tag> struct node_tag<addmod<T, Tag>> { typedef addmod<> type; };
template <class T, class Tag> struct node_tag<submod<T, Tag>> { typedef submod<> type; };
template <class T, class Tag> struct node_tag<mulmod<T, Tag>> { typedef mulmod<> type; };
template <class T, class Tag> struct node_tag<mulmod_shoup<T, Tag>> { typedef mulmod_shoup<> type; };
template <class T, class Tag> struct node_tag<compute_shoup<T, Tag>> { typedef compute_shoup<> type; };
template <class Op, class... Args> expr<typename node_tag<Op>::type, Args...> make_op(Args const &... args) {
  return expr<typename node_tag<Op>::type, Args...>(args...);
}
}  // namespace ops

// shoup(a * b, bprime)  ->  mulmod_shoup(a, b, bprime)   (ops.hpp:266-277)
template <class A0, class A1, class A2>
ops::expr<ops::mulmod_shoup<>, A0, A1, A2> shoup(ops::expr<ops::mulmod<>, A0, A1> const &prod, A2 const &bprime) {
  return ops::expr<ops::mulmod_shoup<>, A0, A1, A2>(std::get<0>(prod.args), std::get<1>(prod.args), bprime);
}
template <class A0> typename std::enable_if<detail::is_operand<A0>::value, ops::expr<ops::compute_shoup<>, A0>>::type compute_shoup(A0 const &a) {
  return ops::expr<ops::compute_shoup<>, A0>(a);
}

namespace detail {
// heap storage for one over-aligned poly (plain `new` does not honour 32-byte alignment in C++11;
// the reference has the same caveat, tests/nfllib_demo_main_func.cpp:40-45)
template <class P> struct aligned_holder {
  P *p;
  aligned_holder() : p(nullptr) {
    void *raw = nullptr;
    if (posix_memalign(&raw, 32, sizeof(P)) != 0) throw std::bad_alloc();
    p = new (raw) P;
  }
  ~aligned_holder() { if (p) { p->~P(); free(p); } }
  aligned_holder(const aligned_holder &) = delete;
  aligned_holder &operator=(const aligned_holder &) = delete;
};
// materialise an operand on the host (used only by the boolean comparisons, which are not on the hot path)
template <class P, class X> struct host_value;
template <class P> struct host_value<P, P> { static P const &get(P const &p, P &) { return p; } };
template <class T, size_t D, size_t M> struct host_value<poly<T, D, M>, poly_p<T, D, M>> {
  static poly<T, D, M> const &get(poly_p<T, D, M> const &p, poly<T, D, M> &) { return p.poly_obj(); }
};
template <class P, class Op, class... A> struct host_value<P, ops::expr<Op, A...>> {
  static P const &get(ops::expr<Op, A...> const &e, P &tmp) { tmp = e; return tmp; }
};
}  // namespace detail

namespace ops {
template <class Op, class... Args> expr<Op, Args...>::operator bool() const {
  static_assert(std::is_same<Op, eqmod>::value || std::is_same<Op, neqmod>::value, "only == and != convert to bool (ops.hpp:81-95)");
  typedef poly_type P;
  detail::aligned_holder<P> ta, tb;
  typedef typename std::tuple_element<0, std::tuple<Args...>>::type X0;
  typedef typename std::tuple_element<1, std::tuple<Args...>>::type X1;
  P const &a = detail::host_value<P, X0>::get(std::get<0>(args), *ta.p);
  P const &b = detail::host_value<P, X1>::get(std::get<1>(args), *tb.p);
  const bool want_equal = std::is_same<Op, eqmod>::value;
  for (size_t i = 0; i < P::degree * P::nmoduli; ++i)
    if ((a.begin()[i] == b.begin()[i]) == want_equal) return true;  // ANY coefficient (the reference's semantics)
  return false;
}
}  // namespace ops

/* High level wrappers (poly.hpp:314-332) and the modulus-size alias (poly.hpp:336-337) */
template <class T, size_t D, size_t M> void sub(poly<T, D, M> &out, poly<T, D, M> const &a, poly<T, D, M> const &b) { out = a - b; }
template <class T, size_t D, size_t M> void add(poly<T, D, M> &out, poly<T, D, M> const &a, poly<T, D, M> const &b) { out = a + b; }
template <class T, size_t D, size_t M> void mul(poly<T, D, M> &out, poly<T, D, M> const &a, poly<T, D, M> const &b) { out = a * b; }
template <class T, size_t Degree, size_t AggregatedModulusBitSize>
using poly_from_modulus = poly<T, Degree, AggregatedModulusBitSize / params<T>::kModulusBitsize>;

/* stream operator (core.hpp:397-421): "{ c0ULL, c1ULL, ... }" over all residues, literal suffix by limb type */
template <class T, size_t D, size_t M> std::ostream &operator<<(std::ostream &os, poly<T, D, M> const &p) {
  const char *term = sizeof(T) == 8 ? "ULL" : sizeof(T) == 4 ? "UL" : "U";
  os << "{ ";
  for (size_t i = 0; i < D * M; ++i) os << (i ? ", " : "") << p.begin()[i] << term;
  return os << " }";
}
template <class T, size_t D, size_t M> std::ostream &operator<<(std::ostream &os, poly_p<T, D, M> const &p) {  // poly_p.hpp:213-217
  return os << p.poly_obj();
}

// ---------------------------------------------------------------------------------------------------------
// Device-resident batches: the throughput API.  `count` polys stay in HBM in exactly the layout of poly[count].
// ---------------------------------------------------------------------------------------------------------
namespace cuda {

// Whole host arrays of polys through the chunked host-buffer ring (nflgpu_host_op: upload, kernel and download of neighbouring chunks overlap): what a loop of
// p[i].ntt_pow_phi() over a contiguous array should be written as.  In place.
template <class P> void ntt_pow_phi(P *polys, size_t count) {
  detail::check(nflgpu_host_op(P::backend_type::get().ctx, 0, polys, polys, nullptr, nullptr, count), "ntt_pow_phi[]");
}
template <class P> void invntt_pow_invphi(P *polys, size_t count) {
  detail::check(nflgpu_host_op(P::backend_type::get().ctx, 1, polys, polys, nullptr, nullptr, count), "invntt_pow_invphi[]");
}
// The same without the final wait (nflgpu_host_op_async): several arrays can be in flight, the uploads of one running under the
// downloads of the previous one; the arrays must stay alive and untouched until host_sync<P>() returns.
template <class P> void ntt_pow_phi_async(P *polys, size_t count) {
  detail::check(nflgpu_host_op_async(P::backend_type::get().ctx, 0, polys, polys, nullptr, nullptr, count), "ntt_pow_phi_async[]");
}
template <class P> void invntt_pow_invphi_async(P *polys, size_t count) {
  detail::check(nflgpu_host_op_async(P::backend_type::get().ctx, 1, polys, polys, nullptr, nullptr, count), "invntt_pow_invphi_async[]");
}
template <class P> void host_sync() { detail::check(nflgpu_host_sync(P::backend_type::get().ctx), "nflgpu_host_sync"); }
// Page-locks a host array of polys for the lifetime of the guard (nflgpu_host_register): the host-buffer calls above then
// DMA it directly instead of staging it through pinned buffers.  Destroy the guard before freeing the array.
template <class P> class pinned_region {
  void *p_;
public:
  pinned_region(P *polys, size_t count) : p_(polys) {
    detail::check(nflgpu_host_register(P::backend_type::get().ctx, polys, count * sizeof(P)), "nflgpu_host_register");
  }
  ~pinned_region() { nflgpu_host_unregister(P::backend_type::get().ctx, p_); }
  pinned_region(const pinned_region &) = delete;
  pinned_region &operator=(const pinned_region &) = delete;
};

template <class P> class batch {
  detail::dev_buf<P> buf_;
  static nflgpu_ctx *ctx() { return P::backend_type::get().ctx; }
  static std::vector<bool> compare(batch const &a, batch const &b, bool want_equal) {
    if (a.size() != b.size()) throw std::runtime_error("nfl::cuda::batch: size mismatch");
    // the flags travel in a (pooled) buffer of whole polys: ceil(count / sizeof(P)) polys hold count bytes
    detail::dev_buf<P> flags((a.size() + sizeof(P) - 1) / sizeof(P));
    uint8_t *f = static_cast<uint8_t *>(flags.p);
    detail::check(want_equal ? nflgpu_any_eq(ctx(), f, a.buf_.p, b.buf_.p, a.size(), nullptr)
                             : nflgpu_any_neq(ctx(), f, a.buf_.p, b.buf_.p, a.size(), nullptr), "nflgpu_any_eq");
    std::vector<uint8_t> host(flags.count * sizeof(P));
    detail::check(nflgpu_download(ctx(), host.data(), flags.p, flags.count, nullptr), "nflgpu_download");
    sync();
    return std::vector<bool>(host.begin(), host.begin() + a.size());
  }

public:
  explicit batch(size_t count) : buf_(count, false) {}
  batch(P const *host, size_t count) : buf_(count, false) { upload(host); }
  size_t size() const { return buf_.count; }
  void *device_ptr() { return buf_.p; }
  const void *device_ptr() const { return buf_.p; }
  void upload(P const *host) { detail::check(nflgpu_upload(ctx(), buf_.p, host, buf_.count, nullptr), "nflgpu_upload"); sync(); }
  void download(P *host) const { detail::check(nflgpu_download(ctx(), host, buf_.p, buf_.count, nullptr), "nflgpu_download"); sync(); }
  static void sync() { detail::check(nflgpu_sync(ctx(), nullptr), "nflgpu_sync"); }

  void ntt_pow_phi() { detail::check(nflgpu_ntt_fwd(ctx(), buf_.p, buf_.p, buf_.count, nullptr), "nflgpu_ntt_fwd"); }
  void invntt_pow_invphi() { detail::check(nflgpu_ntt_inv(ctx(), buf_.p, buf_.p, buf_.count, nullptr), "nflgpu_ntt_inv"); }
  // count successive poly::set(nfl::uniform()) draws, generated in HBM from the Salsa20 stream (key, first_nonce + i):
  // bit-identical to the reference's draws under the same key (core.hpp:150-187, lib/prng/fastrandombytes.cpp:21-34)
  void set_uniform(const uint8_t key[32], uint64_t first_nonce) {
    detail::check(nflgpu_uniform(ctx(), buf_.p, buf_.count, key, first_nonce, nullptr), "nflgpu_uniform");
  }
  void set_non_uniform(uint64_t upper_bound, uint64_t amplifier, const uint8_t key[32], uint64_t first_nonce) {  // core.hpp:190-278
    detail::check(nflgpu_non_uniform(ctx(), buf_.p, buf_.count, upper_bound, amplifier, key, first_nonce, nullptr), "nflgpu_non_uniform");
  }
  void set_hwt(uint32_t hwt, const uint8_t key[32], uint64_t first_nonce) {  // core.hpp:355-392
    detail::check(nflgpu_hwt(ctx(), buf_.p, buf_.count, hwt, key, first_nonce, nullptr), "nflgpu_hwt");
  }
  // count successive poly::set(gaussian(&prng, amplifier)) draws (core.hpp:291-325); returns the nonces they consumed
  template <class in_class, unsigned _lu_depth>
  uint64_t set_gaussian(FastGaussianNoise<in_class, typename P::value_type, _lu_depth> &prng, uint64_t amplifier, const uint8_t key[32],
                        uint64_t first_nonce) {
    uint64_t used = 0;
    detail::check(nflgpu_gaussian_sample(ctx(), prng.handle(ctx()), buf_.p, buf_.count, amplifier, key, first_nonce, &used, nullptr),
                  "nflgpu_gaussian_sample");
    return used;
  }
  void set_zo(uint8_t rho, const uint8_t key[32], uint64_t first_nonce) {  // core.hpp:338-349
    detail::check(nflgpu_zo(ctx(), buf_.p, buf_.count, rho, key, first_nonce, nullptr), "nflgpu_zo");
  }
  // CRT lift (poly::GMP::poly2mpz / mpz2poly, gmp.hpp:183-219) without GMP types: `words` is a device buffer of
  // size() * degree * lift_words() uint64_t, each coefficient as little-endian 64-bit words (mpz_export layout)
  static size_t lift_words() { size_t w = 0; detail::check(nflgpu_lift_words(ctx(), &w), "nflgpu_lift_words"); return w; }
  void poly2words(uint64_t *device_words) const { detail::check(nflgpu_poly2mpz(ctx(), device_words, buf_.p, buf_.count, nullptr), "nflgpu_poly2mpz"); }
  void words2poly(const uint64_t *device_words) { detail::check(nflgpu_mpz2poly(ctx(), buf_.p, device_words, buf_.count, nullptr), "nflgpu_mpz2poly"); }
  // the cyclic transforms underneath (poly::core::ntt / inv_ntt, core.hpp:455-557; what tests/ntt_perfs.cpp times)
  void core_ntt() { detail::check(nflgpu_ntt_raw_fwd(ctx(), buf_.p, buf_.p, buf_.count, nullptr), "nflgpu_ntt_raw_fwd"); }
  void core_inv_ntt() { detail::check(nflgpu_ntt_raw_inv(ctx(), buf_.p, buf_.p, buf_.count, nullptr), "nflgpu_ntt_raw_inv"); }

#define NFLB200_BATCH_BIN(NAME, CALL)                                                                                       \
  void NAME(batch const &a, batch const &b) {                                                                               \
    if (a.size() != size() || b.size() != size()) throw std::runtime_error("nfl::cuda::batch: size mismatch");             \
    detail::check(CALL(ctx(), buf_.p, a.buf_.p, b.buf_.p, buf_.count, nullptr), #CALL);                                    \
  }
  NFLB200_BATCH_BIN(assign_add, nflgpu_add)          // *this = a + b
  NFLB200_BATCH_BIN(assign_sub, nflgpu_sub)          // *this = a - b
  NFLB200_BATCH_BIN(assign_mul, nflgpu_mul)          // *this = a * b
  NFLB200_BATCH_BIN(assign_polymul, nflgpu_polymul)  // *this = invntt(ntt(a) * ntt(b))
#undef NFLB200_BATCH_BIN
  void assign_compute_shoup(batch const &a) { detail::check(nflgpu_compute_shoup(ctx(), buf_.p, a.buf_.p, buf_.count, nullptr), "nflgpu_compute_shoup"); }
  void assign_mul_shoup(batch const &a, batch const &b, batch const &bprime) {
    detail::check(nflgpu_mul_shoup(ctx(), buf_.p, a.buf_.p, b.buf_.p, bprime.buf_.p, buf_.count, nullptr), "nflgpu_mul_shoup");
  }
  // *this = <postfix program over operands> in one pass (nflgpu_eval); e.g. {0,1,2,0x12,0x10} = ops[0] + ops[1]*ops[2]
  void assign_eval(std::vector<batch const *> const &operands, std::vector<uint8_t> const &program) {
    std::vector<const void *> ptrs;
    for (size_t i = 0; i < operands.size(); ++i) {
      if (operands[i]->size() != size()) throw std::runtime_error("nfl::cuda::batch: size mismatch");
      ptrs.push_back(operands[i]->buf_.p);
    }
    detail::check(nflgpu_eval(ctx(), buf_.p, ptrs.data(), ptrs.size(), program.data(), program.size(), buf_.count, nullptr), "nflgpu_eval");
  }
  // operator== / operator!= of every polynomial pair, the reference's ANY-coefficient semantics (ops.hpp:81-117), compared
  // in HBM: result[i] = (a[i] == b[i]) as the reference's expr::operator bool would give it
  static std::vector<bool> any_equal(batch const &a, batch const &b) { return compare(a, b, true); }
  static std::vector<bool> any_different(batch const &a, batch const &b) { return compare(a, b, false); }
  void assign_muladd(batch const &a, batch const &b, batch const &c) {  // *this = a + b * c
    detail::check(nflgpu_muladd(ctx(), buf_.p, a.buf_.p, b.buf_.p, c.buf_.p, buf_.count, nullptr), "nflgpu_muladd");
  }
};

}  // namespace cuda
}  // namespace nfl

#endif  // NFL_B200_HPP

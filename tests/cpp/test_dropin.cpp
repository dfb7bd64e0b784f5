// Drop-in surface test for include/nfl_b200.hpp, compiled as C++11 (the reference's language level).
//
// Mirrors the reference's own tests, but with full-array comparisons instead of its any-equal operator==:
//   tests/test_binary_op.h:10-31 + nfl_add.cpp / nfl_sub.cpp / nfl_mul.cpp   op vs naive per-coefficient lambda
//   tests/poly_p.cpp:52-66                                                   NTT round trip, nested expressions
//   tests/nfllib_demo_main_op.cpp:61-87                                      shoup(a*b, b') vs a*b
//   tests/nfl_eq.cpp / nfl_neq.cpp                                           == / != "any coefficient" semantics
//   tests/poly_set.cpp                                                       set() forms, std::runtime_error on bad sizes
// and dumps results for tests/test_cpp_dropin.py to compare bit-for-bit with the CPU oracle.
//   usage: test_dropin <u64|u32|u16> <in_a.bin> <in_b.bin> <count> <out.bin>
#include <nfl_b200.hpp>

#include <cmath>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <vector>

// the friend proxy exactly as the reference's tests define it (tests/ntt_perfs.cpp:113-131)
namespace nfl { namespace tests {
template <class P> class poly_tests_proxy {
  using value_type = typename P::value_type;
public:
  static inline bool ntt(value_type *x, const value_type *wtab, const value_type *winvtab, value_type const p) { return P::core::ntt(x, wtab, winvtab, p); }
  static inline bool inv_ntt(value_type *x, const value_type *wtab, const value_type *winvtab, value_type invK, value_type const p) {
    return P::core::inv_ntt(x, wtab, winvtab, invK, p);
  }
  static inline value_type *get_omegas(P &p, size_t cm) { return &p.base.omegas[cm][0]; }
  static inline value_type *get_shoupomegas(P &p, size_t cm) { return &p.base.shoupomegas[cm][0]; }
  static inline value_type *get_invomegas(P &p, size_t cm) { return &p.base.invomegas[cm][0]; }
  static inline value_type *get_shoupinvomegas(P &p, size_t cm) { return &p.base.shoupinvomegas[cm][0]; }
  static inline value_type get_invdegree(P &p, size_t cm) { return p.base.invpolyDegree[cm]; }
};
} }

#define REQUIRE(cond) do { if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

template <class P> static P *alloc_polys(size_t n) {  // tests/tools.h:6-17 alloc_aligned
  void *p = nullptr;
  if (posix_memalign(&p, 32, n * sizeof(P)) != 0) return nullptr;
  return new (p) P[n];
}

template <class P> static bool same(P const &a, P const &b) { return std::memcmp(a.begin(), b.begin(), sizeof(P)) == 0; }

template <class P> static int run(const char *fa, const char *fb, size_t count, const char *fout) {
  typedef typename P::value_type T;
  typedef typename P::greater_value_type G;
  static_assert(sizeof(P) == P::degree * P::nmoduli * sizeof(T), "poly must be a plain array (poly.hpp:87-88)");
  static_assert(alignof(P) == 32, "poly must be 32-byte aligned");
  P *a = alloc_polys<P>(count), *b = alloc_polys<P>(count);
  std::ifstream ia(fa, std::ios::binary), ib(fb, std::ios::binary);
  for (size_t i = 0; i < count; ++i) { a[i].deserialize_manually(ia); b[i].deserialize_manually(ib); }
  REQUIRE(ia.good() && ib.good());

  // ---- reference-style unit checks on the first pair --------------------------------------------------
  {
    P &x = a[0], &y = b[0];
    P *tmp = alloc_polys<P>(6);
    P &sum = tmp[0], &dif = tmp[1], &prd = tmp[2], &e = tmp[3], &bs = tmp[4], &ps = tmp[5];
    sum = x + y; dif = x - y; prd = x * y;
    for (size_t cm = 0; cm < P::nmoduli; ++cm) {
      const T p = P::get_modulus(cm);
      for (size_t i = 0; i < P::degree; ++i) {
        e(cm, i) = static_cast<T>((G(x(cm, i)) + y(cm, i)) % p);
      }
    }
    REQUIRE(same(sum, e));                                    // nfl_add.cpp
    for (size_t cm = 0; cm < P::nmoduli; ++cm) for (size_t i = 0; i < P::degree; ++i)
      e(cm, i) = static_cast<T>((G(x(cm, i)) + P::get_modulus(cm) - y(cm, i)) % P::get_modulus(cm));
    REQUIRE(same(dif, e));                                    // nfl_sub.cpp
    for (size_t cm = 0; cm < P::nmoduli; ++cm) for (size_t i = 0; i < P::degree; ++i)
      e(cm, i) = static_cast<T>((G(x(cm, i)) * y(cm, i)) % P::get_modulus(cm));
    REQUIRE(same(prd, e));                                    // nfl_mul.cpp
    nfl::add(sum, x, y); nfl::mul(e, x, y);                   // nfllib_demo_main_func.cpp
    REQUIRE(same(e, prd));
    bs = nfl::compute_shoup(y);
    ps = nfl::shoup(x * y, bs);
    REQUIRE(same(ps, prd));                                   // nfllib_demo_main_op.cpp:61-87
    e = x + y * x;                                            // nested expression, fused muladd (poly_p.cpp:63-66)
    sum = y * x; sum = x + sum;
    REQUIRE(same(e, sum));
    e = (x + y) * (x - y) + y * y;                            // = x*x, 7-token fused program
    sum = x * x;
    REQUIRE(same(e, sum));
    e = x + y * x - nfl::shoup(x * y, bs);                    // fused: add, mul, sub, mul_shoup; = x
    REQUIRE(same(e, x));
    e = ((((x + y) + (y + x)) + ((x - y) + (y - x))) + (((x * y) + (y * x)) - ((x * y) + (x * y)))) - (y + y);  // = 2x
    sum = x + x;
    REQUIRE(same(e, sum));
    P c(x.begin(), x.end(), false);                           // iterator ctor, all residues given
    REQUIRE(same(c, x));
    c.ntt_pow_phi(); REQUIRE(!same(c, x));
    c.invntt_pow_invphi(); REQUIRE(same(c, x));               // poly_p.cpp:52-58, strong form
    // whole host arrays through the host-buffer ring: blocking and asynchronous forms against the per-poly calls
    {
      P *arr = alloc_polys<P>(6), *brr = alloc_polys<P>(6);
      for (int i = 0; i < 6; ++i) { arr[i] = (i & 1) ? y : x; brr[i] = (i & 1) ? x : y; }
      P fx = x, fy = y; fx.ntt_pow_phi(); fy.ntt_pow_phi();
      nfl::cuda::ntt_pow_phi(arr, 6);
      nfl::cuda::ntt_pow_phi_async(brr, 6);
      nfl::cuda::host_sync<P>();
      for (int i = 0; i < 6; ++i) { REQUIRE(same(arr[i], (i & 1) ? fy : fx)); REQUIRE(same(brr[i], (i & 1) ? fx : fy)); }
      nfl::cuda::invntt_pow_invphi_async(arr, 6);
      nfl::cuda::invntt_pow_invphi_async(brr, 6);
      nfl::cuda::host_sync<P>();
      for (int i = 0; i < 6; ++i) { REQUIRE(same(arr[i], (i & 1) ? y : x)); REQUIRE(same(brr[i], (i & 1) ? x : y)); }
      for (int i = 0; i < 6; ++i) { arr[i].~P(); brr[i].~P(); }
      std::free(arr); std::free(brr);
    }
    // == is "any coefficient equal", != is "any coefficient differs" (ops.hpp:81-95)
    REQUIRE(bool(x == x)); REQUIRE(!bool(x != x)); REQUIRE(bool(x != (x + P(1))));
    e = x; e(0, 0) = static_cast<T>((e(0, 0) + 1) % P::get_modulus(0));
    REQUIRE(bool(e == x)); REQUIRE(bool(e != x));             // both true: one differs, the rest are equal
    // set(): constant is reduced per residue and placed in coefficient 0 (core.hpp:76-98)
    P k(static_cast<T>(P::get_modulus(0) + 5));
    REQUIRE(k(0, 0) == 5 && k(0, 1) == 0);
    bool threw = false;
    try { std::vector<T> big(P::degree + 1, 1); P bad(big.begin(), big.end()); (void)bad; } catch (std::runtime_error const &) { threw = true; }
    REQUIRE(threw || P::nmoduli == 1);                        // core.hpp:111-115 (degree+1 == degree*nmoduli only if ... never)
    REQUIRE(P::get_modulus(0) == nfl::params<T>::P[0]);
    // poly_p: same results as poly through the shared handle (tests/poly_p.cpp:12-66, strong comparisons)
    {
      typedef nfl::poly_p<T, P::degree, P::nmoduli> PP;
      PP pa(x.begin(), x.end(), false), pb(y.begin(), y.end(), false);
      PP psum(pa + pb), pdif(pa - pb), pprd(pa * pb);
      sum = x + y; dif = x - y;
      REQUIRE(same(psum.poly_obj(), sum) && same(pdif.poly_obj(), dif) && same(pprd.poly_obj(), prd));
      PP pc(pb);                                              // shares storage
      REQUIRE(&static_cast<PP const &>(pc).poly_obj() == &static_cast<PP const &>(pb).poly_obj());
      pc = {1};                                               // detaches: pb is untouched
      REQUIRE(same(static_cast<PP const &>(pb).poly_obj(), y) && pc(0, 0) == 1 && pc(0, 1) == 0);
      PP pbs = nfl::compute_shoup(pb);
      REQUIRE(same(static_cast<PP const &>(pbs).poly_obj(), bs));
      PP pmul2 = nfl::shoup(pa * pb, pbs);
      REQUIRE(same(static_cast<PP const &>(pmul2).poly_obj(), prd));
      PP pmix = pa + pb * psum;                               // poly_p operands inside a fused expression
      e = x + y * sum;
      REQUIRE(same(static_cast<PP const &>(pmix).poly_obj(), e));
      e = pa + y * psum;                                      // mixed poly / poly_p operands
      REQUIRE(same(static_cast<PP const &>(pmix).poly_obj(), e));
      pa.ntt_pow_phi(); c = x; c.ntt_pow_phi();
      REQUIRE(same(static_cast<PP const &>(pa).poly_obj(), c));
      pa.invntt_pow_invphi();
      REQUIRE(same(static_cast<PP const &>(pa).poly_obj(), x));
      REQUIRE(bool(pa == x) && !bool(pa != x));
    }
    // core::ntt / core::inv_ntt through the reference's friend proxy (tests/ntt_perfs.cpp:122-134,165-171), one residue at
    // a time, against the batch entry points; and the reference-layout omega tables (core.hpp:564-581)
    {
      typedef nfl::tests::poly_tests_proxy<P> proxy;
      nfl::cuda::batch<P> dx(&x, 1);
      dx.core_ntt(); dx.download(&e);
      c = x;
      for (size_t cm = 0; cm < P::nmoduli; ++cm)
        REQUIRE(proxy::ntt(&c(cm, 0), proxy::get_omegas(c, cm), proxy::get_shoupomegas(c, cm), P::get_modulus(cm)));  // always true, core.hpp:531
      REQUIRE(same(c, e));
      dx.core_inv_ntt(); dx.download(&e);
      for (size_t cm = 0; cm < P::nmoduli; ++cm)
        REQUIRE(proxy::inv_ntt(&c(cm, 0), proxy::get_invomegas(c, cm), proxy::get_shoupinvomegas(c, cm), proxy::get_invdegree(c, cm), P::get_modulus(cm)));
      REQUIRE(same(c, e));
      for (size_t cm = 0; cm < P::nmoduli; ++cm) {
        const T p = P::get_modulus(cm);
        const T *w = proxy::get_omegas(c, cm), *ws = proxy::get_shoupomegas(c, cm), *wi = proxy::get_invomegas(c, cm);
        REQUIRE(w[0] == 1 && ws == w + P::degree);
        T t = 1;                                                // omega has order exactly `degree`
        for (size_t i = 0; i < P::degree / 2; ++i) { REQUIRE(w[i] == t); t = static_cast<T>((G(t) * w[1]) % p); }
        REQUIRE(t == p - 1);
        REQUIRE(static_cast<T>((G(w[1]) * wi[1]) % p) == 1);
        REQUIRE(ws[1] == static_cast<T>((G(w[1]) << (8 * sizeof(T))) / p));
        REQUIRE(w[P::degree / 2 + 1] == static_cast<T>((G(w[1]) * w[1]) % p));  // second level: powers of omega^2
        REQUIRE(static_cast<T>((G(proxy::get_invdegree(c, cm)) * (P::degree % p)) % p) == 1);
      }
      bool refused = false;                                     // foreign tables are refused, not silently ignored
      try { proxy::ntt(&c(0, 0), proxy::get_invomegas(c, 0), proxy::get_shoupinvomegas(c, 0), P::get_modulus(0)); } catch (std::runtime_error const &) { refused = true; }
      REQUIRE(refused);
    }
    // random fills come from the device samplers (core.hpp:150-392): ranges and supports
    {
      P u1{nfl::uniform()}, u2{nfl::uniform()};
      REQUIRE(!same(u1, u2) && bool(u1));
      for (size_t cm = 0; cm < P::nmoduli; ++cm) for (size_t i = 0; i < P::degree; ++i) REQUIRE(u1(cm, i) < P::get_modulus(cm));
      P nu{nfl::non_uniform(4, 2)};                             // 2 * (-4, 4), the same centred value in every residue
      P zo{nfl::ZO_dist()};
      P hw{nfl::hwt_dist(P::degree / 8)};
      size_t weight = 0;
      for (size_t i = 0; i < P::degree; ++i) {
        const T p0 = P::get_modulus(0);
        const T v = nu(0, i), z = zo(0, i), h = hw(0, i);
        REQUIRE(v <= 6 || v >= p0 - 6); REQUIRE((v <= 6 ? v : p0 - v) % 2 == 0);
        for (size_t cm = 1; cm < P::nmoduli; ++cm) {
          const T pc = P::get_modulus(cm);
          REQUIRE(v <= 6 ? nu(cm, i) == v : pc - nu(cm, i) == p0 - v);
          REQUIRE((z == 0 && zo(cm, i) == 0) || (z == p0 + 1 && zo(cm, i) == pc + 1) || (z == p0 - 1 && zo(cm, i) == pc - 1));
          REQUIRE((h == 0) == (hw(cm, i) == 0));
        }
        REQUIRE(z == 0 || z == p0 + 1 || z == p0 - 1);             // the reference stores +1 as p + 1 (core.hpp:343)
        REQUIRE(h == 0 || h == p0 + 1 || h == p0 - 1);             // core.hpp:388
        weight += h != 0;
      }
      REQUIRE(weight == P::degree / 8);
      // discrete Gaussian (tests/nfllib_demo_main_op.cpp:141-144,273-283): same small centred value in every residue, within the
      // tail bound (14.4 sigma for these parameters), amplified draws are multiples of the amplifier, and the draws differ
      {
        nfl::FastGaussianNoise<uint8_t, T, 2> fg(3.19, 128, 1 << 14);
        P g1{nfl::gaussian<uint8_t, T, 2>(&fg)}, g2 = nfl::gaussian<uint8_t, T, 2>(&fg, 2);
        REQUIRE(!same(g1, g2));
        double sum = 0, sq = 0;
        for (size_t i = 0; i < P::degree; ++i) {
          const T p0 = P::get_modulus(0);
          const T a = g1(0, i), b = g2(0, i);
          const long va = a <= 64 ? (long)a : -(long)(p0 - a), vb = b <= 128 ? (long)b : -(long)(p0 - b);
          REQUIRE(va >= -47 && va <= 47 && vb >= -94 && vb <= 94 && vb % 2 == 0);
          sum += va; sq += (double)va * va;
          for (size_t cm = 1; cm < P::nmoduli; ++cm) {
            const T pc = P::get_modulus(cm);
            REQUIRE(va >= 0 ? g1(cm, i) == a : pc - g1(cm, i) == p0 - a);
            REQUIRE(vb >= 0 ? g2(cm, i) == b : pc - g2(cm, i) == p0 - b);
          }
        }
        const double mean = sum / P::degree, var = sq / P::degree - mean * mean;
        REQUIRE(std::fabs(mean) < 6 * 3.19 / std::sqrt((double)P::degree) + 1e-9);
        if (P::degree >= 512) REQUIRE(var > 0.6 * 3.19 * 3.19 && var < 1.5 * 3.19 * 3.19);
      }
      bool threw2 = false;
      try { P bad{nfl::non_uniform(P::get_modulus(0))}; (void)bad; } catch (std::runtime_error const &) { threw2 = true; }
      REQUIRE(threw2);                                          // core.hpp:201-206
      std::ostringstream os;                                    // core.hpp:397-421
      P one(1);
      os << one;
      const std::string term = sizeof(T) == 8 ? "ULL" : sizeof(T) == 4 ? "UL" : "U";
      REQUIRE(os.str().substr(0, 3 + term.size() + 2) == "{ 1" + term + ", " && os.str().substr(os.str().size() - term.size() - 3) == "0" + term + " }");
    }
    free(tmp);
  }

  // ---- dump: per poly pair, via the single-poly API and via the batch API --------------------------------
  std::ofstream out(fout, std::ios::binary);
  P *r = alloc_polys<P>(8);
  for (size_t i = 0; i < count; ++i) {
    r[0] = a[i]; r[0].ntt_pow_phi();
    r[1] = a[i]; r[1].invntt_pow_invphi();
    r[2] = a[i] + b[i];
    r[3] = a[i] - b[i];
    r[4] = a[i] * b[i];
    r[5] = nfl::compute_shoup(b[i]);
    r[6] = nfl::shoup(a[i] * b[i], r[5]);
    r[7] = b[i]; r[7].ntt_pow_phi(); r[7] = r[0] * r[7]; r[7].invntt_pow_invphi();   // negacyclic product
    for (int k = 0; k < 8; ++k) r[k].serialize_manually(out);
  }
  {
    typedef nfl::cuda::batch<P> B;
    B da(a, count), db(b, count), t(count), u(count);
    P *h = alloc_polys<P>(count);
    auto dump = [&](B const &x) { x.download(h); for (size_t i = 0; i < count; ++i) h[i].serialize_manually(out); };
    t.assign_add(da, da); t.assign_sub(t, da);                 // t = a
    t.ntt_pow_phi(); dump(t);                                  // fwd(a)
    t.invntt_pow_invphi(); dump(t);                            // a again
    t.core_ntt(); t.core_inv_ntt();                            // = N * a  (core::inv_ntt does not scale, core.hpp:539-557)
    {
      t.download(h);
      for (size_t i = 0; i < count; ++i)
        for (size_t cm = 0; cm < P::nmoduli; ++cm)
          for (size_t j = 0; j < P::degree; j += 97)
            REQUIRE(h[i](cm, j) == static_cast<T>((G(a[i](cm, j)) * (P::degree % P::get_modulus(cm))) % P::get_modulus(cm)));
    }
    t.assign_mul(da, db); dump(t);
    u.assign_compute_shoup(db); t.assign_mul_shoup(da, db, u); dump(t);
    t.assign_muladd(da, db, da);                               // a + b*a
    u.assign_eval({&da, &db}, {0, 1, 0, 0x12, 0x10});          // same through the fused evaluator
    {
      P *h2 = alloc_polys<P>(count);
      t.download(h);
      u.download(h2);
      for (size_t i = 0; i < count; ++i) { REQUIRE(same(h[i], h2[i])); }
      free(h2);
    }
    dump(t);
    t.assign_polymul(da, db); dump(t);
    free(h);
  }
  REQUIRE(out.good());
  free(a); free(b); free(r);
  std::printf("test_dropin ok (%zu polys)\n", count);
  return 0;
}

int main(int argc, char **argv) {
  if (argc != 6) { std::fprintf(stderr, "usage: %s <u64|u32|u16> a.bin b.bin count out.bin\n", argv[0]); return 2; }
  const std::string t = argv[1];
  const size_t count = std::strtoul(argv[4], nullptr, 10);
  try {
    if (t == "u64") return run<nfl::poly<uint64_t, 1024, 4>>(argv[2], argv[3], count, argv[5]);
    if (t == "u32") return run<nfl::poly_from_modulus<uint32_t, 4096, 90>>(argv[2], argv[3], count, argv[5]);  // 3 moduli
    if (t == "u16") return run<nfl::poly<uint16_t, 512, 2>>(argv[2], argv[3], count, argv[5]);
  } catch (std::exception const &e) {
    std::fprintf(stderr, "exception: %s\n", e.what());
    return 3;
  }
  return 2;
}

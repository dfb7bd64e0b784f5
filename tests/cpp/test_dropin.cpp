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

#include <cstdio>
#include <fstream>
#include <vector>

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

// Discrete-Gaussian sampler tables (host, one-time per sampler) — the counterpart of nfl::FastGaussianNoise's constructor
// (include/nfl/prng/FastGaussianNoise.hpp): init() 233-282, precomputeBarrierValues() 285-353, buildLookupTables() 357-476.
//
// The cumulative-distribution ("barrier") table is defined by MPFR arithmetic at a fixed precision, one correctly rounded
// operation after the other; to be bit-identical to the reference it is computed with the same library calls in the same
// order.  MPFR is a build-time dependency of the reference (CMakeLists.txt:34-37); here it is a run-time one: libmpfr.so.6
// and libgmp.so.10 are opened on first use, and a machine without them still runs everything else (and can create a sampler
// from a barrier table computed elsewhere, nflgpu_gaussian_create_from_barriers).  Only plain ABI facts are used: the entry
// points below and the 4-field __mpfr_struct / 3-field __mpz_struct layouts of MPFR 4 / GMP 6.
#include "gaussian.h"
#include "host_common.hpp"

#include <cmath>
#include <cstring>
#include <dlfcn.h>
#include <mutex>

namespace nflgpu {

namespace {

struct mpfr_s { long prec; int sign; long exp; unsigned long *d; };
struct mpz_s { int alloc, size; unsigned long *d; };
enum { RNDN = 0 };

struct Mp {
  void (*init2)(mpfr_s *, long);
  void (*init)(mpfr_s *);
  void (*clear)(mpfr_s *);
  int (*set_d)(mpfr_s *, double, int);
  int (*set_si)(mpfr_s *, long, int);
  int (*set_ui)(mpfr_s *, unsigned long, int);
  int (*set)(mpfr_s *, const mpfr_s *, int);
  int (*sqr)(mpfr_s *, const mpfr_s *, int);
  int (*neg)(mpfr_s *, const mpfr_s *, int);
  int (*exp_)(mpfr_s *, const mpfr_s *, int);
  int (*add)(mpfr_s *, const mpfr_s *, const mpfr_s *, int);
  int (*sub)(mpfr_s *, const mpfr_s *, const mpfr_s *, int);
  int (*mul)(mpfr_s *, const mpfr_s *, const mpfr_s *, int);
  int (*mul_ui)(mpfr_s *, const mpfr_s *, unsigned long, int);
  int (*sub_ui)(mpfr_s *, const mpfr_s *, unsigned long, int);
  int (*ui_div)(mpfr_s *, unsigned long, const mpfr_s *, int);
  int (*pow_ui)(mpfr_s *, const mpfr_s *, unsigned long, int);
  int (*get_z)(mpz_s *, const mpfr_s *, int);
  void (*free_cache)(void);
  void (*z_init2)(mpz_s *, unsigned long);
  void (*z_clear)(mpz_s *);
  size_t (*z_sizeinbase)(const mpz_s *, int);
  void *(*z_export)(void *, size_t *, int, size_t, int, size_t, const mpz_s *);
  bool ok = false;
};

const Mp &mp() {
  static Mp m;
  static std::once_flag once;
  std::call_once(once, [] {
    void *hf = dlopen("libmpfr.so.6", RTLD_NOW | RTLD_LOCAL), *hz = dlopen("libgmp.so.10", RTLD_NOW | RTLD_LOCAL);
    if (!hf || !hz) return;
    bool all = true;
#define SYM(h, field, name) all = all && ((*(void **)(&m.field) = dlsym(h, name)) != nullptr)
    SYM(hf, init2, "mpfr_init2"); SYM(hf, init, "mpfr_init"); SYM(hf, clear, "mpfr_clear"); SYM(hf, set_d, "mpfr_set_d");
    SYM(hf, set_si, "mpfr_set_si"); SYM(hf, set_ui, "mpfr_set_ui"); SYM(hf, set, "mpfr_set"); SYM(hf, sqr, "mpfr_sqr");
    SYM(hf, neg, "mpfr_neg"); SYM(hf, exp_, "mpfr_exp"); SYM(hf, add, "mpfr_add"); SYM(hf, sub, "mpfr_sub"); SYM(hf, mul, "mpfr_mul");
    SYM(hf, mul_ui, "mpfr_mul_ui"); SYM(hf, sub_ui, "mpfr_sub_ui"); SYM(hf, ui_div, "mpfr_ui_div"); SYM(hf, pow_ui, "mpfr_pow_ui");
    SYM(hf, get_z, "mpfr_get_z"); SYM(hf, free_cache, "mpfr_free_cache");
    SYM(hz, z_init2, "__gmpz_init2"); SYM(hz, z_clear, "__gmpz_clear"); SYM(hz, z_sizeinbase, "__gmpz_sizeinbase");
    SYM(hz, z_export, "__gmpz_export");
#undef SYM
    m.ok = all;
  });
  return m;
}

// the tail bound t solves t^2 - 2 ln t - 1 - 2k ln 2 = 0 (FastGaussianNoise.hpp:118-158, the non-Boost solver: the
// reference's build never defines BOOST_RAPHSON); same double expressions, three digits
double tail_bound_of(double k, double start) {
  double guess = start;
  for (unsigned counter = 0; counter < (1u << 15); ++counter) {
    const double f = guess * guess - 2 * std::log(guess) - 1 - 2 * k * std::log(2), df = 2 * guess - 2 / guess;
    const double delta = f / df;
    guess -= delta;
    if (std::fabs(delta) / std::fabs(guess) < std::pow(10.0, -3)) break;
  }
  while (0.95 * guess * 0.95 * guess - 2 * std::log(0.95 * guess) - 1 - 2 * k * std::log(2) >= 0) guess *= 0.95;
  return guess;
}

}  // namespace

bool gaussian_runtime_available() { return mp().ok; }

// init() + precomputeBarrierValues(): fills t->nb, wp, bit_precision, tail_bound, rounded_center, barriers
int gaussian_compute_barriers(double sigma, unsigned security, unsigned samples, double center, int in_bytes, GaussianTable *t) {
  const Mp &m = mp();
  if (!m.ok) { set_error("Gaussian tables need the MPFR/GMP runtimes (libmpfr.so.6, libgmp.so.10), which could not be loaded"); return -2; }
  if (!(sigma > 0) || samples == 0 || (in_bytes != 1 && in_bytes != 2)) { set_error("bad Gaussian parameters"); return -1; }
  // FastGaussianNoise.hpp:252-270
  const double k = security + 1 + std::ceil(std::log(samples) / std::log(2));
  const double tail = tail_bound_of(k, std::sqrt(1 + 2 * k * std::log(2)));
  const double epsi = k + std::log2(2 * tail * sigma);
  unsigned bit_precision = (unsigned)std::ceil(epsi);
  const unsigned wp = (unsigned)std::ceil(bit_precision / (8.0 * in_bytes));
  bit_precision = wp * 8 * in_bytes;
  const unsigned nb = (unsigned)(1 + 2 * std::ceil(tail * sigma));
  t->nb = nb; t->wp = wp; t->bit_precision = bit_precision; t->tail_bound = tail; t->in_bytes = in_bytes;
  t->rounded_center = (long)std::round(center);  // :176
  t->barriers.assign((size_t)nb * wp * in_bytes, 0);

  mpfr_s cs, ctr, sum, tmp, tmp2;
  m.init2(&cs, bit_precision); m.init2(&sum, bit_precision); m.init2(&tmp, bit_precision); m.init2(&tmp2, bit_precision);
  m.init(&ctr);                                   // default precision, :175
  m.set_d(&ctr, center, RNDN);
  m.set_d(&cs, sigma, RNDN); m.sqr(&cs, &cs, RNDN); m.mul_ui(&cs, &cs, 2, RNDN); m.ui_div(&cs, 1, &cs, RNDN);  // 1 / (2 sigma^2), :273-277
  std::vector<mpfr_s> acc(nb);
  m.set_ui(&sum, 0, RNDN);
  for (unsigned i = 0; i < nb; ++i) {  // :306-326
    m.init2(&acc[i], bit_precision);
    m.set_si(&tmp2, t->rounded_center + (long)i - ((long)nb - 1) / 2, RNDN);
    // nn_gaussian_law :632-639: exp(-(x - c)^2 / (2 sigma^2)), every step rounded to bit_precision
    m.sub(&tmp, &tmp2, &ctr, RNDN); m.sqr(&tmp, &tmp, RNDN); m.neg(&tmp, &tmp, RNDN); m.mul(&tmp, &tmp, &cs, RNDN); m.exp_(&tmp, &tmp, RNDN);
    if (i == 0) m.set(&acc[0], &tmp, RNDN);
    else m.add(&acc[i], &acc[i - 1], &tmp, RNDN);
    m.add(&sum, &sum, &tmp, RNDN);
  }
  // scale = (2^bit_precision - 1) / sum, :329-333
  m.ui_div(&sum, 1, &sum, RNDN);
  m.set_ui(&tmp, 2, RNDN); m.pow_ui(&tmp, &tmp, bit_precision, RNDN); m.sub_ui(&tmp, &tmp, 1, RNDN);
  m.mul(&sum, &sum, &tmp, RNDN);
  mpz_s z;
  m.z_init2(&z, bit_precision);
  int rc = 0;
  for (unsigned i = 0; i < nb; ++i) {  // :336-352: round to an integer, store most significant word first, right aligned
    m.mul(&acc[i], &acc[i], &sum, RNDN);
    m.get_z(&z, &acc[i], RNDN);
    const unsigned words = (unsigned)std::ceil((float)m.z_sizeinbase(&z, 256) / in_bytes);
    if (words > wp) { rc = -3; set_error("Gaussian barrier does not fit its precision"); }
    else m.z_export(t->barriers.data() + ((size_t)i * wp + (wp - words)) * in_bytes, nullptr, 1, in_bytes, 0, 0, &z);
    m.clear(&acc[i]);
  }
  m.z_clear(&z);
  m.clear(&cs); m.clear(&ctr); m.clear(&sum); m.clear(&tmp); m.clear(&tmp2);
  m.free_cache();
  return rc;
}

static inline uint32_t bword(const GaussianTable &t, size_t b, size_t j) {
  const unsigned char *r = t.barriers.data() + (b * t.wp + j) * t.in_bytes;
  return t.in_bytes == 1 ? r[0] : (uint32_t)r[0] | ((uint32_t)r[1] << 8);
}

// buildLookupTables(): first-word table, and for depth 2 one second-word table per flagged first word.  The linked lists of
// barrier pointers of the reference are ranges here: barriers are consumed in ascending order.
int gaussian_build_luts(GaussianTable *t, int depth) {
  if (!((t->in_bytes == 1 && (depth == 1 || depth == 2)) || (t->in_bytes == 2 && depth == 1))) {
    set_error("Gaussian look-up: supported shapes are (uint8_t, depth 1 | 2) and (uint16_t, depth 1)");
    return -1;
  }
  if (t->nb == 0 || t->wp < (unsigned)depth || (size_t)t->wp * t->in_bytes > GAUSS_MAX_ROW_BYTES) { set_error("bad Gaussian barrier table"); return -1; }
  const size_t lu = t->in_bytes == 1 ? 256 : 65536, nb = t->nb;
  t->depth = depth; t->lu_size = (unsigned)lu; t->flag_ctr1 = t->flag_ctr2 = 0;
  t->lut.assign(lu, GaussLutEntry{0, -1, 0, 0});
  size_t i1 = 0, b = 0;
  long val = -((long)nb - 1) / 2 + t->rounded_center;
  const long last = ((long)nb - 1) / 2 + t->rounded_center;
  while (val <= last && i1 < lu) {
    while (i1 < bword(*t, b, 0) && i1 < lu) t->lut[i1++].val = (int32_t)val;
    GaussLutEntry first{(int32_t)val, 0, 0, 0};
    ++t->flag_ctr1;
    if (depth == 1) {
      first.bstart = (uint32_t)b;
      ++b; ++val;
      while (b < nb && i1 == bword(*t, b, 0)) { ++b; ++val; }
      first.bcount = (uint32_t)b - first.bstart;
    } else {
      first.sub = (int32_t)(t->lut.size() / lu);  // second-level table number (>= 1: table 0 is the first level)
      const size_t base = t->lut.size();
      t->lut.resize(base + lu, GaussLutEntry{0, -1, 0, 0});
      for (size_t i2 = 0; i2 < lu; ++i2) {
        GaussLutEntry &e = t->lut[base + i2];
        if (b >= nb || i1 < bword(*t, b, 0) || i2 < bword(*t, b, 1)) {
          e.val = (int32_t)val;
        } else if (i1 == bword(*t, b, 0) && i2 == bword(*t, b, 1)) {
          e.val = (int32_t)val; e.sub = 0; e.bstart = (uint32_t)b;
          ++t->flag_ctr2;
          ++b; ++val;
          while (b < nb && i1 == bword(*t, b, 0) && i2 == bword(*t, b, 1)) { ++b; ++val; }
          e.bcount = (uint32_t)b - e.bstart;
        }
      }
    }
    t->lut[i1++] = first;
  }
  return 0;
}

// getNoise() :489-501: look-up words drawn per refill for `degree` outputs — float arithmetic, exactly as the reference writes it
uint64_t gaussian_words_per_fill(const GaussianTable &t, uint64_t degree) {
  const unsigned lu = t.lu_size, f1 = t.flag_ctr1, f2 = t.flag_ctr2, wp = t.wp;
  float mult;
  if (t.depth == 1) mult = 1.05 * ((float)(lu - f1) / (float)lu) + wp * ((float)f1 / lu);
  else mult = 1.05 * ((float)(lu - f1) / (float)lu) + 2.0 * ((float)f1 / (float)lu) + wp * ((float)f2 / ((float)lu * lu));
  return (uint64_t)(degree * mult);
}

}  // namespace nflgpu

/* Declaration-only stand-in for <mpfr.h> (TEST INFRASTRUCTURE, oracle build only); see gmp.h here.
 * Prototypes and the __mpfr_struct layout follow the public MPFR 4 manual / ABI of libmpfr.so.6 (present in the image
 * without its development header), so the reference's Gaussian sampler (prng/FastGaussianNoise.hpp) links against the
 * installed runtime.  mpfr_init_set* are macros in the real header and are provided inline here the same way;
 * mpfr_out_str is exported as __gmpfr_out_str. */
#ifndef NFLB200_ORACLE_SHIM_MPFR_H
#define NFLB200_ORACLE_SHIM_MPFR_H
#include <gmp.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef long mpfr_prec_t;
typedef long mpfr_exp_t;
typedef struct { mpfr_prec_t _mpfr_prec; int _mpfr_sign; mpfr_exp_t _mpfr_exp; mp_limb_t *_mpfr_d; } __mpfr_struct;
typedef __mpfr_struct mpfr_t[1];
typedef __mpfr_struct *mpfr_ptr;
typedef const __mpfr_struct *mpfr_srcptr;
typedef enum { MPFR_RNDN = 0, MPFR_RNDZ, MPFR_RNDU, MPFR_RNDD, MPFR_RNDA } mpfr_rnd_t;

void mpfr_init2(mpfr_ptr, mpfr_prec_t);
void mpfr_inits2(mpfr_prec_t, mpfr_ptr, ...);
void mpfr_clear(mpfr_ptr);
void mpfr_clears(mpfr_ptr, ...);
void mpfr_free_cache(void);
void mpfr_init(mpfr_ptr);
int mpfr_set(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_set_d(mpfr_ptr, double, mpfr_rnd_t);
int mpfr_set_si(mpfr_ptr, long, mpfr_rnd_t);
int mpfr_set_ui(mpfr_ptr, unsigned long, mpfr_rnd_t);
double mpfr_get_d(mpfr_srcptr, mpfr_rnd_t);
int mpfr_get_z(mpz_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_add(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_sub(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_sub_ui(mpfr_ptr, mpfr_srcptr, unsigned long, mpfr_rnd_t);
int mpfr_mul(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_mul_ui(mpfr_ptr, mpfr_srcptr, unsigned long, mpfr_rnd_t);
int mpfr_ui_div(mpfr_ptr, unsigned long, mpfr_srcptr, mpfr_rnd_t);
int mpfr_sqr(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_neg(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_exp(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_pow_ui(mpfr_ptr, mpfr_srcptr, unsigned long, mpfr_rnd_t);
#define mpfr_out_str __gmpfr_out_str
size_t mpfr_out_str(FILE *, int, size_t, mpfr_srcptr, mpfr_rnd_t);
int mpfr_set_d(mpfr_ptr, double, mpfr_rnd_t);
static inline int mpfr_init_set(mpfr_ptr x, mpfr_srcptr y, mpfr_rnd_t r) { mpfr_init(x); return mpfr_set(x, y, r); }
static inline int mpfr_init_set_d(mpfr_ptr x, double d, mpfr_rnd_t r) { mpfr_init(x); return mpfr_set_d(x, d, r); }
#ifdef __cplusplus
}
#endif
#endif

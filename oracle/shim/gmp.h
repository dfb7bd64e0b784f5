/* Declaration-only stand-in for <gmp.h> (TEST INFRASTRUCTURE, oracle build only).
 *
 * The image ships libgmp.so.10 but no headers.  The reference's poly.hpp:33 and
 * prng/FastGaussianNoise.hpp:15 include <gmp.h>/<gmpxx.h>/<mpfr.h> unconditionally, but nothing on the
 * NTT / pointwise hot path instantiates a GMP- or MPFR-using member, so prototypes
 * are enough for `#include <nfl.hpp>` to compile; no mpz_ or mpfr_ symbol is ever referenced at link time.
 * The prototypes follow the public GMP 6 manual. */
#ifndef NFLB200_ORACLE_SHIM_GMP_H
#define NFLB200_ORACLE_SHIM_GMP_H
#include <stddef.h>
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef unsigned long mp_limb_t;
typedef unsigned long mp_bitcnt_t;
typedef struct { int _mp_alloc; int _mp_size; mp_limb_t *_mp_d; } __mpz_struct;
typedef __mpz_struct mpz_t[1];
typedef __mpz_struct *mpz_ptr;
typedef const __mpz_struct *mpz_srcptr;

/* The runtime library (libgmp.so.10, present in the image) exports these functions under the __gmpz_ prefix; the real
 * <gmp.h> maps the public names with macros exactly like this.  With them the reference's CRT-lifting code
 * (include/nfl/gmp.hpp) links against the installed runtime although its development headers are absent. */
#define mpz_init2 __gmpz_init2
#define mpz_inits __gmpz_inits
#define mpz_clear __gmpz_clear
#define mpz_clears __gmpz_clears
#define mpz_init_set_ui __gmpz_init_set_ui
#define mpz_set_ui __gmpz_set_ui
#define mpz_mul __gmpz_mul
#define mpz_mul_ui __gmpz_mul_ui
#define mpz_addmul_ui __gmpz_addmul_ui
#define mpz_sub __gmpz_sub
#define mpz_submul __gmpz_submul
#define mpz_divexact __gmpz_divexact
#define mpz_tdiv_q __gmpz_tdiv_q
#define mpz_tdiv_q_2exp __gmpz_tdiv_q_2exp
#define mpz_ui_pow_ui __gmpz_ui_pow_ui
#define mpz_invert __gmpz_invert
#define mpz_cmp __gmpz_cmp
#define mpz_fdiv_ui __gmpz_fdiv_ui
#define mpz_sizeinbase __gmpz_sizeinbase
#define mpz_out_str __gmpz_out_str
#define mpz_export __gmpz_export
#define mpz_import __gmpz_import
#define mpz_init __gmpz_init

void mpz_init(mpz_ptr);
void mpz_import(mpz_ptr, size_t, int, size_t, int, size_t, const void *);
void mpz_init2(mpz_ptr, mp_bitcnt_t);
void mpz_inits(mpz_ptr, ...);
void mpz_clear(mpz_ptr);
void mpz_clears(mpz_ptr, ...);
void mpz_init_set_ui(mpz_ptr, unsigned long);
void mpz_set_ui(mpz_ptr, unsigned long);
void mpz_mul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_mul_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_addmul_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_sub(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_submul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_divexact(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_tdiv_q(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_tdiv_q_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void mpz_ui_pow_ui(mpz_ptr, unsigned long, unsigned long);
int mpz_invert(mpz_ptr, mpz_srcptr, mpz_srcptr);
int mpz_cmp(mpz_srcptr, mpz_srcptr);
unsigned long mpz_fdiv_ui(mpz_srcptr, unsigned long);
size_t mpz_sizeinbase(mpz_srcptr, int);
size_t mpz_out_str(FILE *, int, mpz_srcptr);
void *mpz_export(void *, size_t *, int, size_t, int, size_t, mpz_srcptr);
#ifdef __cplusplus
}
#endif
#endif

/* Declaration-only stand-in for <gmpxx.h> (TEST INFRASTRUCTURE, oracle build only); see gmp.h here. */
#ifndef NFLB200_ORACLE_SHIM_GMPXX_H
#define NFLB200_ORACLE_SHIM_GMPXX_H
#include <gmp.h>
class mpz_class {
  mpz_t mp;
public:
  mpz_class();
  mpz_class(const mpz_class &);
  mpz_class(mpz_srcptr);
  mpz_class(unsigned long);
  ~mpz_class();
  mpz_class &operator=(const mpz_class &);
  mpz_srcptr get_mpz_t() const { return mp; }
  mpz_ptr get_mpz_t() { return mp; }
};
#endif

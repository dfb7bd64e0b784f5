// TEST INFRASTRUCTURE — dispatcher over the parallel-built parts of ref_harness.cpp (see there).
#include <cstddef>
extern "C" {
#define DECL(k) int nflref_run_part##k(int, int, size_t, size_t, void *, const void *, const void *, const void *, size_t, int);
DECL(0) DECL(1) DECL(2) DECL(3) DECL(4) DECL(5)
int nflref_run(int op, int limb_bits, size_t degree, size_t nmoduli, void *out, const void *a, const void *b,
               const void *c, size_t batch, int threads) {
  int rc;
#define TRY(k) rc = nflref_run_part##k(op, limb_bits, degree, nmoduli, out, a, b, c, batch, threads); if (rc != -1) return rc;
  TRY(0) TRY(1) TRY(2) TRY(3) TRY(4) TRY(5)
  return -1;
}
}

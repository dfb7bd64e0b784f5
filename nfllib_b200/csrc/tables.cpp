// Twiddle-table construction (host, one-time per context) — the B200 counterpart of
// poly::core::initialize() / prep_wtab() (core.hpp:625-686, 564-581).
//
// The reference stores omega-power tables per DIF stage plus separate phi^i / N^-1 phi^-i twist tables
// (poly.hpp:228-237).  The kernels use the merged formulation instead: Cooley-Tukey butterflies with
// psi = phi powers in bit-reversed order produce ntt_pow_phi()'s output directly (natural input,
// bit-reversed evaluation order), and Gentleman-Sande butterflies with psi^-1 powers consume that order
// and produce invntt_pow_invphi()'s output, with N^-1 folded into the last stage.  Only the mathematical
// function is kept; the tables are laid out for conflict-free per-pass reads (ntt_plan.h).
#include "host_common.hpp"
#include "ntt_plan.h"

namespace nflgpu {

typedef unsigned __int128 u128;

static inline uint64_t mm(uint64_t a, uint64_t b, uint64_t p) { return (uint64_t)(((u128)a * b) % p); }
static inline uint64_t shoup_of(uint64_t w, uint64_t p, int limb_bits) {
  return (uint64_t)((((u128)w) << limb_bits) / p);  // floor(w * 2^w / p), core.hpp:575,653
}
static inline size_t bitrev(size_t v, int bits) {
  size_t r = 0;
  for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
  return r;
}

void build_residue_tables(int limb_bits, int word_bits, size_t N, uint64_t p, uint64_t root, uint64_t kmax,
                          ResidueTables *out, bool raw) {
  int n = 0;
  while (((size_t)1 << n) < N) ++n;
  // core.hpp:640-645: psi = root^(kmax / N) by repeated squaring; a primitive 2N-th root of unity
  uint64_t psi = root;
  for (uint64_t k = kmax; k > N; k >>= 1) psi = mm(psi, psi, p);
  uint64_t ipsi = invmod64(psi, p);  // = psi^(2N-1), core.hpp:661
  // core.hpp:664-665: N^-1
  uint64_t ninv = invmod64((uint64_t)N % p, p);

  std::vector<uint64_t> pw(N), ipw(N);  // psi^i, psi^-i
  pw[0] = ipw[0] = 1;
  for (size_t i = 1; i < N; ++i) { pw[i] = mm(pw[i - 1], psi, p); ipw[i] = mm(ipw[i - 1], ipsi, p); }

  const bool fold = plan_fold(n, word_bits);  // N^-1 carried by the twiddles (ntt_plan.h plan_fold)
  const uint64_t scale = raw ? 1 : ninv;      // (core::inv_ntt leaves the scaling to the twist that follows it, core.hpp:539-557,613)
  out->fwd_w.assign(N, 0); out->fwd_ws.assign(N, 0);
  out->inv_w.assign((size_t)plan_inv_entries(n, word_bits), 0); out->inv_ws.assign((size_t)plan_inv_entries(n, word_bits), 0);
  const int np = plan_npass(n, word_bits);
  for (int i = 0; i < np && n > 0; ++i) {
    const int r = plan_r(n, word_bits, i), s0 = plan_s0(n, word_bits, i);
    const size_t G = (size_t)1 << s0, off = plan_off(n, word_bits, i);
    for (int q = 0; q < r; ++q)
      for (size_t kk = 0; kk < ((size_t)1 << q); ++kk)
        for (size_t g = 0; g < G; ++g) {
          size_t e_idx = ((size_t)1 << q) - 1 + kk;
          size_t k = ((size_t)1 << (s0 + q)) + (g << q) + kk;  // index into the bit-reversed power table
          // merged (negacyclic) tables: psi^bitrev_n(k) = psi^(t + 2t*bitrev_s(i)), t = N/2^(s+1), i = group index;
          // raw (cyclic, core::ntt / core::inv_ntt without the phi twist): the same without the psi^t factor
          size_t ex = raw ? (bitrev(k, n) - ((size_t)1 << (n - 1 - (s0 + q)))) : bitrev(k, n);
          size_t at = off + e_idx * G + g;
          uint64_t w = pw[ex], iw = ipw[ex];
          const uint64_t iws = mm(iw, scale, p);
          out->fwd_w[at] = w; out->fwd_ws[at] = shoup_of(w, p, limb_bits);
          if (fold) {
            // plain, except the stage that pairs position bit 0 (forward stage n-1), whose butterflies all take the scaled twiddle;
            // the other stages of that pass (the last of the plan) have their scaled copies behind the table, same [e_idx][g] order
            const uint64_t m = (s0 + q == n - 1) ? iws : iw;
            out->inv_w[at] = m; out->inv_ws[at] = shoup_of(m, p, limb_bits);
            if (i == np - 1 && s0 + q != n - 1) {
              out->inv_w[N + e_idx * G + g] = iws; out->inv_ws[N + e_idx * G + g] = shoup_of(iws, p, limb_bits);
            }
          } else {
            const uint64_t m = (k == 1) ? iws : iw;  // the last inverse stage also scales its difference output by N^-1
            out->inv_w[at] = m; out->inv_ws[at] = shoup_of(m, p, limb_bits);
          }
        }
  }
  // slot N-1: N^-1 (1 for the raw transform) -- applied to the (U+V) output of the last inverse stage, or with folded tables to register 0 of every thread after the first inverse pass
  out->inv_w[N - 1] = scale;
  out->inv_ws[N - 1] = shoup_of(scale, p, limb_bits);
}

}  // namespace nflgpu

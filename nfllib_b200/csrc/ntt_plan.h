// Pass plan of the multi-pass register-radix NTT, shared by the host table builder and the kernels.
//
// A transform of size N = 2^n is done in NP passes; in each pass a thread holds E = 2^e coefficients in
// registers and runs r_i <= e butterfly stages on them; between passes the coefficients of one
// (residue, polynomial) unit are exchanged through (padded) shared memory.
//
// Position bits of a coefficient index are numbered n-1 (MSB) .. 0.  Forward stage s (s = 0 .. n-1) pairs
// positions that differ in bit n-1-s.  Pass i covers stages [s0_i, s0_i + r_i), i.e. position bits
// [hi_i - 1 .. hi_i - r_i] with hi_i = n - s0_i; the thread's register index k is the window of e position
// bits [hi_i - 1 .. c_i], c_i = hi_i - e (for a short first pass the low e - r_0 bits of k are independent
// columns).  The remaining bits form the thread id inside the unit: tid = (g << c_i) | l with g the bits
// above the window (the butterfly group) and l the bits below it.
//
// Twiddle table of one residue, one direction: N entries {w, shoup(w)}.  Pass i owns
// (2^r_i - 1) * 2^s0_i consecutive entries starting at off_i, indexed [e_idx][g] with
// e_idx = 2^q - 1 + kk for stage s0_i + q and kk the top q bits of k:  entry = psi^( +-bitrev_n(2^(s0_i+q) +
// (g << q) + kk) ).  Entry N-1 is unused by the forward table and holds N^-1 in the inverse table.
#ifndef NFLGPU_NTT_PLAN_H
#define NFLGPU_NTT_PLAN_H

#if defined(__CUDACC__)
#define NFLGPU_HD __host__ __device__
#else
#define NFLGPU_HD
#endif

namespace nflgpu {

// Largest radix exponent per thread.  64-bit words: 32 coefficients = 64 data registers (16 for N = 1024, where the
// smaller register footprint doubles the resident warps and measured 10 % faster, profiles/r01b_*); 32-bit words: 64
// coefficients up to N = 2048, 32 above (the shapes (2,5,5) .. (5,5,5) beat (6,6) / (1,6,6) on the N = 4096 config).
// NFLGPU_EMAX64 / NFLGPU_EMAX32 override the table (tools/variants.sh experiments).
NFLGPU_HD constexpr int plan_emax(int n, int word_bits) {
#ifdef NFLGPU_EMAX64
  if (word_bits == 64) return NFLGPU_EMAX64;
#endif
#ifdef NFLGPU_EMAX32
  if (word_bits != 64) return NFLGPU_EMAX32;
#endif
  return word_bits == 64 ? (n == 10 ? 4 : 5) : (n >= 12 ? 5 : 6);
}
NFLGPU_HD constexpr int plan_npass(int n, int word_bits) { return (n + plan_emax(n, word_bits) - 1) / plan_emax(n, word_bits); }
NFLGPU_HD constexpr int plan_e(int n, int word_bits) {
#ifdef NFLGPU_FORCE_E  // experiment builds (one size at a time): e.g. 5 gives N = 4096 the shape (2,5,5) instead of (4,4,4)
  return NFLGPU_FORCE_E;
#else
  return (n + plan_npass(n, word_bits) - 1) / plan_npass(n, word_bits);
#endif
}
// stages in pass i (only the first pass may be short)
NFLGPU_HD constexpr int plan_r(int n, int word_bits, int i) {
  return i == 0 ? n - plan_e(n, word_bits) * (plan_npass(n, word_bits) - 1) : plan_e(n, word_bits);
}
// first stage of pass i
NFLGPU_HD constexpr int plan_s0(int n, int word_bits, int i) {
  return i == 0 ? 0 : plan_r(n, word_bits, 0) + (i - 1) * plan_e(n, word_bits);
}
NFLGPU_HD constexpr int plan_hi(int n, int word_bits, int i) { return n - plan_s0(n, word_bits, i); }
NFLGPU_HD constexpr int plan_c(int n, int word_bits, int i) { return plan_hi(n, word_bits, i) - plan_e(n, word_bits); }
// offset of pass i in the twiddle table: sum over earlier stages of 2^s = 2^s0 - 1
NFLGPU_HD constexpr int plan_off(int n, int word_bits, int i) { return (1 << plan_s0(n, word_bits, i)) - 1; }

// Transforms too large for one shared-memory tile (64-bit words: N > 2^14) run their first `split` passes as
// global-memory passes (one kernel each, registers <-> HBM, no exchange needed because a pass's E coefficients live in
// one thread); after them the unit has decomposed into 2^s0 independent sub-blocks of 2^hi words, each of which the
// tile kernel finishes exactly like a small transform (its twiddles are indexed by the sub-block number).
NFLGPU_HD constexpr int plan_tile_log(int word_bits) { return word_bits == 64 ? 14 : 15; }
NFLGPU_HD constexpr int plan_split(int n, int word_bits) {
  int s = 0;
  while (plan_hi(n, word_bits, s) > plan_tile_log(word_bits)) ++s;
  return s;
}

}  // namespace nflgpu
#endif

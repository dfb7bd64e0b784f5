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

// Inverse transform, 64-bit words: N^-1 is folded into the twiddles instead of being applied by one extra Shoup multiplication per
// butterfly of the last stage (N/2 multiplications per unit, 1 / (log2 N + 1) of all of them).  The inverse pairs position bits
// 0, 1, .. n-1 in that order, and its first pass (NP-1) holds the low e position bits of a thread's coefficients in the register
// index.  If, inside that pass, a coefficient carries the factor N^-1 exactly when its bits below the current one are not all
// zero, both inputs of a butterfly agree, the sum output keeps the invariant for free and the difference output has to pick the
// factor up only when those low bits ARE all zero -- there the butterfly uses the twiddle w * N^-1 instead of w, a compile-time
// choice per register index.  After the pass only register 0 of every thread lacks the factor: one multiplication per thread
// (N / E per unit instead of N / 2), and the remaining passes and the last stage work on scaled values with plain twiddles.
// The inverse table of such a shape has N + N/2 entries per residue: [0, N) as before except that the stage pairing bit 0 (every
// butterfly of it is a "low bits all zero" one) holds the scaled values and no entry is pre-multiplied for the last stage;
// [N, N + N/2) = the entries of pass NP-1's other stages times N^-1, in the same [e_idx][g] order.  (32-bit words keep the extra
// multiplication: three multiply instructions there, no more than the additional table reads.)
// Measured on B200 against the old form (profiles/r02_variants.log block 14): bit-exact everywhere, and 3-4 % fewer multiply-pipe cycles at
// an unchanged instruction count in the SASS, but only the cluster kernel of N = 2^15 gains (inverse 216.7 -> 212.9 us, -1.8 %); the
// tile kernels lose -- N = 1024 +1.6 %, N = 4096 +0.3 %, N = 8192 +8 %, N = 16384 +6.6 % -- because the additional twiddle registers
// of the first pass spill at their register budgets (72 / 128).  So it is on for N = 2^15 only; -DNFLGPU_FOLD=1 / 0 forces it on for
// every 64-bit size / off.
NFLGPU_HD constexpr bool plan_fold(int n, int word_bits) {
#ifdef NFLGPU_FOLD
  return NFLGPU_FOLD != 0 && word_bits == 64 && n >= 1;
#else
  return word_bits == 64 && n == 15;
#endif
}
// entries per residue of the inverse twiddle table
NFLGPU_HD constexpr int plan_inv_entries(int n, int word_bits) { return plan_fold(n, word_bits) ? (1 << n) + (1 << (n - 1)) : (1 << n); }

}  // namespace nflgpu
#endif

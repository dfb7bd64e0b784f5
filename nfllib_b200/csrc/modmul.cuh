// Exact (canonical) per-limb modular multiply and Shoup-word computation, shared by the pointwise kernels and the
// fused forward-NTT * operand epilogue.  Reference: ops::mulmod ops.hpp:184-219, ops::compute_shoup ops.hpp:165-177.
#ifndef NFLGPU_MODMUL_CUH
#define NFLGPU_MODMUL_CUH

#include "modarith.cuh"
#include "pointwise.h"  // enum PwOp

namespace nflgpu {

// ---- per-limb exact modular multiply (any correct x*y mod p is bit-identical) -----------

template <int LB> struct PW;

template <> struct PW<64> {
  typedef uint64_t Word;
  typedef uint64_t Store;
  static constexpr int VEC = 2;
  // ops.hpp:201-219: res = x*y (128 bit); q = Pn*hi(res) + (res << 2); r = lo(res) - hi(q)*p; r -= p if r >= p.
  // (2^128 / p = 4*2^64 + Pn for NFLlib's 62-bit moduli.)
  static NFLGPU_DEVFN Word mulmod(Word x, Word y, Word p, uint64_t pn) {
    const Word lo = x * y, hi = umul64hi_hd(x, y);
    const Word a_lo = pn * hi, a_hi = umul64hi_hd(pn, hi);
    const Word b_lo = lo << 2, b_hi = (hi << 2) | (lo >> 62);
    const Word s_lo = a_lo + b_lo;
    const Word q_hi = a_hi + b_hi + (s_lo < a_lo ? 1 : 0);
    Word r = lo - q_hi * p;
    return csub(r, p);
  }
  // floor(x * 2^64 / p) for x < p:  estimate with mu = 4*2^64 + pn, then at most two corrections
  static NFLGPU_DEVFN Word shoup_of(Word x, Word p, uint64_t pn) {
    Word q = (x << 2) + umul64hi_hd(x, pn);
    Word rem = (Word)0 - q * p;  // x*2^64 - q*p, exact because it is < 3p < 2^64
    if (rem >= p) { rem -= p; ++q; }
    if (rem >= p) { rem -= p; ++q; }
    return q;
  }
  static NFLGPU_DEVFN Word mulhi(Word a, Word b) { return umul64hi_hd(a, b); }
};

template <> struct PW<32> {
  typedef uint32_t Word;
  typedef uint32_t Store;
  static constexpr int VEC = 4;
  // (x*y) % p (ops.hpp:184-197) through Barrett with mu = floor(2^64 / p): q is exact or one short
  static NFLGPU_DEVFN Word mulmod(Word x, Word y, Word p, uint64_t mu) {
    const uint64_t res = (uint64_t)x * y;
    const uint64_t q = umul64hi_hd(res, mu);
    Word r = (Word)res - (Word)q * p;
    return csub(r, p);
  }
  static NFLGPU_DEVFN Word shoup_of(Word x, Word p, uint64_t mu) {
    const uint64_t num = (uint64_t)x << 32;
    uint64_t q = umul64hi_hd(num, mu);
    Word rem = (Word)0 - (Word)q * p;  // num - q*p < 2p
    if (rem >= p) ++q;
    return (Word)q;
  }
  static NFLGPU_DEVFN Word mulhi(Word a, Word b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (Word)(((uint64_t)a * b) >> 32);
#endif
  }
};

template <> struct PW<16> {
  typedef uint32_t Word;
  typedef uint16_t Store;
  static constexpr int VEC = 8;
  static NFLGPU_DEVFN Word mulmod(Word x, Word y, Word p, uint64_t) { return (x * y) % p; }
  static NFLGPU_DEVFN Word shoup_of(Word x, Word p, uint64_t) { return (x << 16) / p; }
  static NFLGPU_DEVFN Word mulhi(Word a, Word b) { return (a * b) >> 16; }
};

// ---- the coefficient-wise functors of the pointwise kernels (pointwise.cu) and of the host simulation -------------------
template <int LB, int OP> struct Functor {
  typedef typename PW<LB>::Word Word;
  static NFLGPU_DEVFN Word apply(Word a, Word b, Word c, Word d, Word p, uint64_t k) {
    if (OP == PW_ADD) return csub(a + b, p);                               // ops.hpp:132-133
    if (OP == PW_SUB) return csub(a + (p - b), p);                         // ops.hpp:149
    if (OP == PW_MUL) return PW<LB>::mulmod(a, b, p, k);                   // ops.hpp:184-219
    if (OP == PW_MUL_SHOUP) {                                              // ops.hpp:231-241
      const Word q = PW<LB>::mulhi(a, c);
      return csub((Word)(a * b - q * p), p);
    }
    if (OP == PW_COMPUTE_SHOUP) {                                          // ops.hpp:170-176
      while (a >= p) a -= p;
      return PW<LB>::shoup_of(a, p, k);
    }
    if (OP == PW_MULADD) return csub(a + PW<LB>::mulmod(b, c, p, k), p);   // a + b*c
    if (OP == PW_MULADD_SHOUP) {                                           // a + shoup(b*c, c') (opt/ops.hpp:56-78)
      const Word q = PW<LB>::mulhi(b, d);
      return csub(a + csub((Word)(b * c - q * p), p), p);
    }
    return 0;
  }
};

}  // namespace nflgpu
#endif

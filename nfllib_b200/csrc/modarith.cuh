// Device modular arithmetic for the three NFLlib limb types (sm_100a, integer pipes only).
//
// Shoup / Harvey lazy arithmetic as in the reference butterflies (algos.hpp:27-42) and functors
// (ops.hpp:124-242), restated for the GPU: 64-bit limbs use IMAD.WIDE chains through __umul64hi, 32-bit limbs
// use __umulhi, 16-bit limbs are widened to 32-bit words (a 4p < 2^16 value times a 16-bit Shoup word fits).
#ifndef NFLGPU_MODARITH_CUH
#define NFLGPU_MODARITH_CUH

#include <cstdint>
#include <cuda_runtime.h>

// The arithmetic and the butterfly networks also compile for the host (plain C in place of the PTX), so that a CPU test can
// run the very code of the kernels, pass by pass, against the oracle (tests/cpp/engine_sim.cu) -- a test device, not a
// product path: nothing in libnflgpu.so calls these functions on the host.
#define NFLGPU_DEVFN __host__ __device__ __forceinline__

namespace nflgpu {

static NFLGPU_DEVFN uint64_t umul64hi_hd(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

template <int LIMB_BITS> struct Arith;

template <> struct Arith<64> {
  typedef uint64_t Word;   // register / shared-memory word
  typedef uint64_t Store;  // global-memory limb
  typedef ulonglong2 TW;   // {w, shoup(w)}
  static constexpr int WORD_BITS = 64;
  static NFLGPU_DEVFN Word tw_w(const TW &t) { return t.x; }
  static NFLGPU_DEVFN Word tw_ws(const TW &t) { return t.y; }
  // (measured, profiles/r01_integer_pipe_model.md: integer code on sm_100 costs ~2 issue cycles per IMAD / IMAD.WIDE /
  //  IADD3-class instruction and about twice that per IMAD.HI, so the four IMAD.WIDE of __umul64hi beat any formulation
  //  built on IMAD.HI.)
  static NFLGPU_DEVFN Word mulhi(Word a, Word b) {
    return umul64hi_hd(a, b);
  }
  // y*w - floor(y*ws / 2^64)*p  in [0, 2p) for any 64-bit y  (algos.hpp:37-38).  `np` is -p mod 2^64 (kept opaque
  // to the optimiser by the caller) so the whole right-hand side is one multiply-accumulate chain, no subtraction.
  static NFLGPU_DEVFN Word mul_shoup_lazy(Word y, Word w, Word ws, Word np) {
    const Word q = mulhi(y, ws);
#if (!defined(NFLGPU_MAD_CHAIN) || NFLGPU_MAD_CHAIN) && defined(__CUDA_ARCH__)
    // lo64(y*w + q*np) as two IMAD.WIDE on one 64-bit accumulator and four IMAD on its high word: six instructions, no
    // separate additions of the cross products
    uint32_t y0 = (uint32_t)y, y1 = (uint32_t)(y >> 32), w0 = (uint32_t)w, w1 = (uint32_t)(w >> 32);
    uint32_t q0 = (uint32_t)q, q1 = (uint32_t)(q >> 32), n0 = (uint32_t)np, n1 = (uint32_t)(np >> 32);
    Word r;
    asm("{\n\t.reg .b64 acc;\n\t.reg .b32 lo, hi;\n\t"
        "mul.wide.u32 acc, %1, %3;\n\t"
        "mad.wide.u32 acc, %5, %7, acc;\n\t"
        "mov.b64 {lo, hi}, acc;\n\t"
        "mad.lo.u32 hi, %2, %3, hi;\n\t"
        "mad.lo.u32 hi, %1, %4, hi;\n\t"
        "mad.lo.u32 hi, %6, %7, hi;\n\t"
        "mad.lo.u32 hi, %5, %8, hi;\n\t"
        "mov.b64 %0, {lo, hi};\n\t}"
        : "=l"(r) : "r"(y0), "r"(y1), "r"(w0), "r"(w1), "r"(q0), "r"(q1), "r"(n0), "r"(n1));
    return r;
#else
    return y * w + q * np;
#endif
  }
};

template <> struct Arith<32> {
  typedef uint32_t Word;
  typedef uint32_t Store;
  typedef uint2 TW;
  static constexpr int WORD_BITS = 32;
  static NFLGPU_DEVFN Word tw_w(const TW &t) { return t.x; }
  static NFLGPU_DEVFN Word tw_ws(const TW &t) { return t.y; }
  // hi32(a*b) through IMAD.WIDE instead of the IMAD.HI (about two issue slots) that __umulhi compiles to
  static NFLGPU_DEVFN Word mulhi(Word a, Word b) {
#ifdef __CUDA_ARCH__
    Word hi;
    asm("{\n\t.reg .b64 t;\n\t.reg .b32 lo;\n\tmul.wide.u32 t, %1, %2;\n\tmov.b64 {lo, %0}, t;\n\t}" : "=r"(hi) : "r"(a), "r"(b));
    return hi;
#else
    return (Word)(((uint64_t)a * b) >> 32);
#endif
  }
  static NFLGPU_DEVFN Word mul_shoup_lazy(Word y, Word w, Word ws, Word np) {
    const Word q = mulhi(y, ws);
    return y * w + q * np;
  }
};

template <> struct Arith<16> {
  typedef uint32_t Word;  // 16-bit limbs are computed in 32-bit words
  typedef uint16_t Store;
  typedef uint2 TW;
  static constexpr int WORD_BITS = 32;
  static NFLGPU_DEVFN Word tw_w(const TW &t) { return t.x; }
  static NFLGPU_DEVFN Word tw_ws(const TW &t) { return t.y; }
  // y < 2^16 (lazy values stay below 4p < 2^16), ws < 2^16: the products are exact in 32 bits
  static NFLGPU_DEVFN Word mul_shoup_lazy(Word y, Word w, Word ws, Word np) {
    const Word q = (y * ws) >> 16;
    return y * w + q * np;  // np = -p mod 2^32; the true value y*w - q*p < 2p < 2^15 survives the wrap
  }
  static NFLGPU_DEVFN Word mulhi(Word a, Word b) { return (a * b) >> 16; }
};

// x - (x >= m ? m : 0)
template <class W> static NFLGPU_DEVFN W csub(W x, W m) { return x >= m ? x - m : x; }
// (the pointwise kernels keep this compare-and-select form: they are HBM-bound and measured 5-10 % slower with the
//  VIADDMNMX form used by the NTT butterflies below)
// Same, for the lazy-range reductions where m <= 2^(w-1) and x < 2m: the sign of x - m decides, which costs one
// compare on the high word instead of a two-instruction 64-bit unsigned compare.
static NFLGPU_DEVFN uint64_t csub_lazy(uint64_t x, uint64_t m) {
  const int64_t t = (int64_t)(x - m);
  return t < 0 ? x : (uint64_t)t;
}
// 32-bit words: the wrapped difference is larger than x exactly when x < m, so an unsigned minimum does the select
// (IADD3 + VIMNMX.U32: two ALU instructions instead of three)
static NFLGPU_DEVFN uint32_t csub_lazy(uint32_t x, uint32_t m) {
#ifdef __CUDA_ARCH__
  return min(x, x - m);
#else
  return x < m ? x : x - m;
#endif
}
// 64-bit words, "top-bit" lazy reduction: x + n2p (n2p = -2p mod 2^64) when bit 63 of x is set, else x.  One compare on the
// high word that does not wait for a subtraction, and a predicated add; ptxas emits four instructions (ISETP, predicated
// IADD3, IADD3.X, SEL) where the compare-after-subtract form takes five.
// Forward butterflies keep every value anywhere in [0, 2^64): with X' = csub_top(X) < max(2^63, 2^64 - 2p) and T < 2p,
// X' + T and X' - T + 2p stay below 2^64 for any p <= 2^62.
static NFLGPU_DEVFN uint64_t csub_top(uint64_t x, uint64_t n2p) {
#ifdef __CUDA_ARCH__
  uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
  asm("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %1, 0;\n\t@p add.cc.u32 %0, %0, %2;\n\t@p addc.u32 %1, %1, %3;\n\t}"
      : "+r"(lo), "+r"(hi) : "r"((uint32_t)n2p), "r"((uint32_t)(n2p >> 32)));
  return ((uint64_t)hi << 32) | lo;
#else
  return (x >> 63) ? x + n2p : x;
#endif
}
static NFLGPU_DEVFN uint32_t csub_top(uint32_t x, uint32_t) { return x; }  // (32-bit words keep csub_lazy)
// The same reduction written as a select (experiment build -DNFLGPU_LAZY64=2): five instructions like csub_lazy, carries in
// freely allocated predicates, and the compare still independent of the add — for the kernels where the predicated form loses.
static NFLGPU_DEVFN uint64_t csub_top_select(uint64_t x, uint64_t n2p) { return (int64_t)x < 0 ? x + n2p : x; }
static NFLGPU_DEVFN uint32_t csub_top_select(uint32_t x, uint32_t) { return x; }
// a - b + c as one three-input add with two carries (IADD3 + IADD3.X); written in PTX so that the front end
// does not reassociate the sum into two separate 64-bit additions
static NFLGPU_DEVFN uint64_t subadd(uint64_t a, uint64_t b, uint64_t c) {
#ifdef __CUDA_ARCH__
  uint64_t r;
  asm("{\n\t.reg .u64 t;\n\tsub.u64 t, %1, %2;\n\tadd.u64 %0, t, %3;\n\t}" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
#else
  return a - b + c;
#endif
}
static NFLGPU_DEVFN uint32_t subadd(uint32_t a, uint32_t b, uint32_t c) { return a - b + c; }
// any 64-bit representative -> [0, p) for p = 2^62 - c:  x = q*2^62 + r  ==>  x = r + q*c (mod p), r + q*c < p + 4c
static NFLGPU_DEVFN uint64_t canon_full(uint64_t x, uint64_t p, uint64_t c) {
  const uint64_t v = (x & 0x3fffffffffffffffull) + (uint64_t)(uint32_t)(x >> 62) * c;
  return csub_lazy(v, p);
}
// -p mod 2^w, hidden from constant propagation so products with it are not rewritten back into subtractions
static NFLGPU_DEVFN uint64_t opaque_neg(uint64_t p) { uint64_t n = (uint64_t)0 - p; asm("" : "+l"(n)); return n; }
static NFLGPU_DEVFN uint32_t opaque_neg(uint32_t p) { uint32_t n = (uint32_t)0 - p; asm("" : "+r"(n)); return n; }

}  // namespace nflgpu
#endif

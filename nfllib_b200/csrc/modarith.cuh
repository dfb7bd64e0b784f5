// Device modular arithmetic for the three NFLlib limb types (sm_100a, integer pipes only).
//
// Shoup / Harvey lazy arithmetic as in the reference butterflies (algos.hpp:27-42) and functors
// (ops.hpp:124-242), restated for the GPU: 64-bit limbs use IMAD.WIDE chains through __umul64hi, 32-bit limbs
// use __umulhi, 16-bit limbs are widened to 32-bit words (a 4p < 2^16 value times a 16-bit Shoup word fits).
#ifndef NFLGPU_MODARITH_CUH
#define NFLGPU_MODARITH_CUH

#include <cstdint>
#include <cuda_runtime.h>

namespace nflgpu {

template <int LIMB_BITS> struct Arith;

template <> struct Arith<64> {
  typedef uint64_t Word;   // register / shared-memory word
  typedef uint64_t Store;  // global-memory limb
  typedef ulonglong2 TW;   // {w, shoup(w)}
  static constexpr int WORD_BITS = 64;
  static __device__ __forceinline__ Word tw_w(const TW &t) { return t.x; }
  static __device__ __forceinline__ Word tw_ws(const TW &t) { return t.y; }
  // y*w - floor(y*ws / 2^64)*p  in [0, 2p) for any 64-bit y  (algos.hpp:37-38)
  static __device__ __forceinline__ Word mul_shoup_lazy(Word y, Word w, Word ws, Word p) {
    Word q = __umul64hi(y, ws);
    return y * w - q * p;
  }
  static __device__ __forceinline__ Word mulhi(Word a, Word b) { return __umul64hi(a, b); }
};

template <> struct Arith<32> {
  typedef uint32_t Word;
  typedef uint32_t Store;
  typedef uint2 TW;
  static constexpr int WORD_BITS = 32;
  static __device__ __forceinline__ Word tw_w(const TW &t) { return t.x; }
  static __device__ __forceinline__ Word tw_ws(const TW &t) { return t.y; }
  static __device__ __forceinline__ Word mul_shoup_lazy(Word y, Word w, Word ws, Word p) {
    Word q = __umulhi(y, ws);
    return y * w - q * p;
  }
  static __device__ __forceinline__ Word mulhi(Word a, Word b) { return __umulhi(a, b); }
};

template <> struct Arith<16> {
  typedef uint32_t Word;  // 16-bit limbs are computed in 32-bit words
  typedef uint16_t Store;
  typedef uint2 TW;
  static constexpr int WORD_BITS = 32;
  static __device__ __forceinline__ Word tw_w(const TW &t) { return t.x; }
  static __device__ __forceinline__ Word tw_ws(const TW &t) { return t.y; }
  // y < 2^16 (lazy values stay below 4p < 2^16), ws < 2^16: the products are exact in 32 bits
  static __device__ __forceinline__ Word mul_shoup_lazy(Word y, Word w, Word ws, Word p) {
    Word q = (y * ws) >> 16;
    return y * w - q * p;
  }
  static __device__ __forceinline__ Word mulhi(Word a, Word b) { return (a * b) >> 16; }
};

// x - (x >= m ? m : 0)
template <class W> static __device__ __forceinline__ W csub(W x, W m) { return x >= m ? x - m : x; }

}  // namespace nflgpu
#endif

// 16-byte vector loads / stores of limbs for the coefficient-wise kernels (pointwise.cu, eval_static.cu).
#ifndef NFLGPU_VECIO_CUH
#define NFLGPU_VECIO_CUH
#include <cstdint>
#include <cuda_runtime.h>

namespace nflgpu {

template <int LB> struct VecIO;
template <> struct VecIO<64> {
  static __device__ __forceinline__ void load(uint64_t (&w)[2], const uint64_t *g) {
    const ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2 *>(g));
    w[0] = t.x; w[1] = t.y;
  }
  static __device__ __forceinline__ void store(uint64_t *g, const uint64_t (&w)[2]) {
    *reinterpret_cast<ulonglong2 *>(g) = make_ulonglong2(w[0], w[1]);
  }
};
template <> struct VecIO<32> {
  static __device__ __forceinline__ void load(uint32_t (&w)[4], const uint32_t *g) {
    const uint4 t = __ldg(reinterpret_cast<const uint4 *>(g));
    w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
  }
  static __device__ __forceinline__ void store(uint32_t *g, const uint32_t (&w)[4]) {
    *reinterpret_cast<uint4 *>(g) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <> struct VecIO<16> {
  static __device__ __forceinline__ void load(uint32_t (&w)[8], const uint16_t *g) {
    const uint4 t = __ldg(reinterpret_cast<const uint4 *>(g));
    const uint32_t v[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { w[2 * i] = v[i] & 0xffffu; w[2 * i + 1] = v[i] >> 16; }
  }
  static __device__ __forceinline__ void store(uint16_t *g, const uint32_t (&w)[8]) {
    *reinterpret_cast<uint4 *>(g) = make_uint4(w[0] | (w[1] << 16), w[2] | (w[3] << 16), w[4] | (w[5] << 16), w[6] | (w[7] << 16));
  }
};

}  // namespace nflgpu
#endif

// 32-bit limbs, forward direction fused with the coefficient-wise product (nflgpu_polymul): degrees 2^3 .. 2^15 (params<uint32_t>::kMaxPolyDegree = 32768).
#include "ntt_launch.cuh"
namespace nflgpu {
cudaError_t launch_ntt_u32_fwdmul(int log2_degree, const NttLaunch &l, int device, int num_sms, cudaStream_t stream) {
  switch (log2_degree) {
    NFLGPU_NTT_CASE(32, 3, 2) NFLGPU_NTT_CASE(32, 4, 2) NFLGPU_NTT_CASE(32, 5, 2) NFLGPU_NTT_CASE(32, 6, 2)
    NFLGPU_NTT_CASE(32, 7, 2) NFLGPU_NTT_CASE(32, 8, 2) NFLGPU_NTT_CASE(32, 9, 2) NFLGPU_NTT_CASE(32, 10, 2)
    NFLGPU_NTT_CASE(32, 11, 2) NFLGPU_NTT_CASE(32, 12, 2) NFLGPU_NTT_CASE(32, 13, 2) NFLGPU_NTT_CASE(32, 14, 2)
    NFLGPU_NTT_CASE(32, 15, 2)
  }
  return cudaErrorInvalidValue;
}
}  // namespace nflgpu

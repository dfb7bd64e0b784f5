// 16-bit limbs, forward direction fused with the coefficient-wise product (nflgpu_polymul): degrees 2^4 .. 2^9 (params<uint16_t>::kMaxPolyDegree = 512).
#include "ntt_launch.cuh"
namespace nflgpu {
cudaError_t launch_ntt_u16_fwdmul(int log2_degree, const NttLaunch &l, int device, int num_sms, cudaStream_t stream) {
  switch (log2_degree) {
    NFLGPU_NTT_CASE(16, 4, 2) NFLGPU_NTT_CASE(16, 5, 2) NFLGPU_NTT_CASE(16, 6, 2) NFLGPU_NTT_CASE(16, 7, 2)
    NFLGPU_NTT_CASE(16, 8, 2) NFLGPU_NTT_CASE(16, 9, 2)
  }
  return cudaErrorInvalidValue;
}
}  // namespace nflgpu

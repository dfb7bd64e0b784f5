// 32-bit limbs, inv direction: degrees 2^3 .. 2^15 (params<uint32_t>::kMaxPolyDegree = 32768).
#include "ntt_launch.cuh"
namespace nflgpu {
cudaError_t launch_ntt_u32_inv(int log2_degree, const NttLaunch &l, int device, int num_sms, cudaStream_t stream) {
  switch (log2_degree) {
    NFLGPU_NTT_CASE(32, 3, 1) NFLGPU_NTT_CASE(32, 4, 1) NFLGPU_NTT_CASE(32, 5, 1) NFLGPU_NTT_CASE(32, 6, 1)
    NFLGPU_NTT_CASE(32, 7, 1) NFLGPU_NTT_CASE(32, 8, 1) NFLGPU_NTT_CASE(32, 9, 1) NFLGPU_NTT_CASE(32, 10, 1)
    NFLGPU_NTT_CASE(32, 11, 1) NFLGPU_NTT_CASE(32, 12, 1) NFLGPU_NTT_CASE(32, 13, 1) NFLGPU_NTT_CASE(32, 14, 1)
    NFLGPU_NTT_CASE(32, 15, 1)
  }
  return cudaErrorInvalidValue;
}
}  // namespace nflgpu

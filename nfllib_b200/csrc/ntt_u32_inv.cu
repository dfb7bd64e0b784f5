// 32-bit limbs, inv direction: degrees 2^3 .. 2^15 (params<uint32_t>::kMaxPolyDegree = 32768).
#include "ntt_launch.cuh"
namespace nflgpu {
cudaError_t launch_ntt_u32_inv(int log2_degree, const NttLaunch &l, int device, int num_sms, cudaStream_t stream) {
  switch (log2_degree) {
    NFLGPU_NTT_CASE(32, 3, true) NFLGPU_NTT_CASE(32, 4, true) NFLGPU_NTT_CASE(32, 5, true) NFLGPU_NTT_CASE(32, 6, true)
    NFLGPU_NTT_CASE(32, 7, true) NFLGPU_NTT_CASE(32, 8, true) NFLGPU_NTT_CASE(32, 9, true) NFLGPU_NTT_CASE(32, 10, true)
    NFLGPU_NTT_CASE(32, 11, true) NFLGPU_NTT_CASE(32, 12, true) NFLGPU_NTT_CASE(32, 13, true) NFLGPU_NTT_CASE(32, 14, true)
    NFLGPU_NTT_CASE(32, 15, true)
  }
  return cudaErrorInvalidValue;
}
}  // namespace nflgpu

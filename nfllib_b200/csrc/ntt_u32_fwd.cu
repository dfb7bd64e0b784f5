// 32-bit limbs, fwd direction: degrees 2^3 .. 2^15 (params<uint32_t>::kMaxPolyDegree = 32768).
#include "ntt_launch.cuh"
namespace nflgpu {
cudaError_t launch_ntt_u32_fwd(int log2_degree, const NttLaunch &l, int device, int num_sms, cudaStream_t stream) {
  switch (log2_degree) {
    NFLGPU_NTT_CASE(32, 3, 0) NFLGPU_NTT_CASE(32, 4, 0) NFLGPU_NTT_CASE(32, 5, 0) NFLGPU_NTT_CASE(32, 6, 0)
    NFLGPU_NTT_CASE(32, 7, 0) NFLGPU_NTT_CASE(32, 8, 0) NFLGPU_NTT_CASE(32, 9, 0) NFLGPU_NTT_CASE(32, 10, 0)
    NFLGPU_NTT_CASE(32, 11, 0) NFLGPU_NTT_CASE(32, 12, 0) NFLGPU_NTT_CASE(32, 13, 0) NFLGPU_NTT_CASE(32, 14, 0)
    NFLGPU_NTT_CASE(32, 15, 0)
  }
  return cudaErrorInvalidValue;
}
}  // namespace nflgpu

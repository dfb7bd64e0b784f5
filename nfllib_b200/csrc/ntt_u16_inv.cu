// 16-bit limbs, inv direction: degrees 2^4 .. 2^9 (params<uint16_t>::kMaxPolyDegree = 512).
#include "ntt_launch.cuh"
namespace nflgpu {
cudaError_t launch_ntt_u16_inv(int log2_degree, const NttLaunch &l, int device, int num_sms, cudaStream_t stream) {
  switch (log2_degree) {
    NFLGPU_NTT_CASE(16, 4, true) NFLGPU_NTT_CASE(16, 5, true) NFLGPU_NTT_CASE(16, 6, true) NFLGPU_NTT_CASE(16, 7, true)
    NFLGPU_NTT_CASE(16, 8, true) NFLGPU_NTT_CASE(16, 9, true)
  }
  return cudaErrorInvalidValue;
}
}  // namespace nflgpu

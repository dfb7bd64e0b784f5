// 64-bit limbs, fwd direction: degrees 2^2 .. 2^14 (a 2^15 tile of 64-bit words exceeds 227 KB of shared memory).
#include "ntt_launch.cuh"
namespace nflgpu {
cudaError_t launch_ntt_u64_fwd(int log2_degree, const NttLaunch &l, int device, int num_sms, cudaStream_t stream) {
  switch (log2_degree) {
    NFLGPU_NTT_CASE(64, 2, false) NFLGPU_NTT_CASE(64, 3, false) NFLGPU_NTT_CASE(64, 4, false) NFLGPU_NTT_CASE(64, 5, false)
    NFLGPU_NTT_CASE(64, 6, false) NFLGPU_NTT_CASE(64, 7, false) NFLGPU_NTT_CASE(64, 8, false) NFLGPU_NTT_CASE(64, 9, false)
    NFLGPU_NTT_CASE(64, 10, false) NFLGPU_NTT_CASE(64, 11, false) NFLGPU_NTT_CASE(64, 12, false) NFLGPU_NTT_CASE(64, 13, false)
    NFLGPU_NTT_CASE(64, 14, false)
  }
  return cudaErrorInvalidValue;
}
}  // namespace nflgpu

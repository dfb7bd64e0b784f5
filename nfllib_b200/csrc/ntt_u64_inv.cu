// 64-bit limbs, inv direction: degrees 2^2 .. 2^14 (a 2^15 tile of 64-bit words exceeds 227 KB of shared memory).
#include "ntt_launch.cuh"
namespace nflgpu {
cudaError_t launch_ntt_u64_inv(int log2_degree, const NttLaunch &l, int device, int num_sms, cudaStream_t stream) {
  switch (log2_degree) {
    NFLGPU_NTT_CASE(64, 2, true) NFLGPU_NTT_CASE(64, 3, true) NFLGPU_NTT_CASE(64, 4, true) NFLGPU_NTT_CASE(64, 5, true)
    NFLGPU_NTT_CASE(64, 6, true) NFLGPU_NTT_CASE(64, 7, true) NFLGPU_NTT_CASE(64, 8, true) NFLGPU_NTT_CASE(64, 9, true)
    NFLGPU_NTT_CASE(64, 10, true) NFLGPU_NTT_CASE(64, 11, true) NFLGPU_NTT_CASE(64, 12, true) NFLGPU_NTT_CASE(64, 13, true)
    NFLGPU_NTT_CASE(64, 14, true)
  }
  return cudaErrorInvalidValue;
}
}  // namespace nflgpu

// 64-bit limbs, forward direction fused with the coefficient-wise product (nflgpu_polymul): degrees 2^2 .. 2^14 (a 2^15 tile of 64-bit words exceeds 227 KB of shared memory).
#include "ntt_launch.cuh"
namespace nflgpu {
cudaError_t launch_ntt_u64_fwdmul(int log2_degree, const NttLaunch &l, int device, int num_sms, cudaStream_t stream) {
  switch (log2_degree) {
    NFLGPU_NTT_CASE(64, 2, 2) NFLGPU_NTT_CASE(64, 3, 2) NFLGPU_NTT_CASE(64, 4, 2) NFLGPU_NTT_CASE(64, 5, 2)
    NFLGPU_NTT_CASE(64, 6, 2) NFLGPU_NTT_CASE(64, 7, 2) NFLGPU_NTT_CASE(64, 8, 2) NFLGPU_NTT_CASE(64, 9, 2)
    NFLGPU_NTT_CASE(64, 10, 2) NFLGPU_NTT_CASE(64, 11, 2) NFLGPU_NTT_CASE(64, 12, 2) NFLGPU_NTT_CASE(64, 13, 2)
    NFLGPU_NTT_CASE(64, 14, 2)
  }
  return cudaErrorInvalidValue;
}
}  // namespace nflgpu

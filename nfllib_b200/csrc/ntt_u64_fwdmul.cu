// 64-bit limbs, forward direction fused with the coefficient-wise product (nflgpu_polymul): degrees 2^2 .. 2^20 = params<uint64_t>::kMaxPolyDegree (above 2^14 the leading passes run as global-memory kernels, ntt_plan.h).
#include "ntt_launch.cuh"
namespace nflgpu {
cudaError_t launch_ntt_u64_fwdmul(int log2_degree, const NttLaunch &l, int device, int num_sms, cudaStream_t stream) {
  switch (log2_degree) {
    NFLGPU_NTT_CASE(64, 2, 2) NFLGPU_NTT_CASE(64, 3, 2) NFLGPU_NTT_CASE(64, 4, 2) NFLGPU_NTT_CASE(64, 5, 2)
    NFLGPU_NTT_CASE(64, 6, 2) NFLGPU_NTT_CASE(64, 7, 2) NFLGPU_NTT_CASE(64, 8, 2) NFLGPU_NTT_CASE(64, 9, 2)
    NFLGPU_NTT_CASE(64, 10, 2) NFLGPU_NTT_CASE(64, 11, 2) NFLGPU_NTT_CASE(64, 12, 2) NFLGPU_NTT_CASE(64, 13, 2)
    NFLGPU_NTT_CASE(64, 14, 2)
    NFLGPU_NTT_CASE(64, 15, 2) NFLGPU_NTT_CASE(64, 16, 2) NFLGPU_NTT_CASE(64, 17, 2) NFLGPU_NTT_CASE(64, 18, 2)
    NFLGPU_NTT_CASE(64, 19, 2) NFLGPU_NTT_CASE(64, 20, 2)
  }
  return cudaErrorInvalidValue;
}
}  // namespace nflgpu

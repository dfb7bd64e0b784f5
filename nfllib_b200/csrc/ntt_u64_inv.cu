// 64-bit limbs, inv direction: degrees 2^2 .. 2^20 = params<uint64_t>::kMaxPolyDegree (above 2^14 the leading passes run as global-memory kernels, ntt_plan.h).
#include "ntt_launch.cuh"
namespace nflgpu {
cudaError_t launch_ntt_u64_inv(int log2_degree, const NttLaunch &l, int device, int num_sms, cudaStream_t stream) {
  switch (log2_degree) {
    NFLGPU_NTT_CASE(64, 2, 1) NFLGPU_NTT_CASE(64, 3, 1) NFLGPU_NTT_CASE(64, 4, 1) NFLGPU_NTT_CASE(64, 5, 1)
    NFLGPU_NTT_CASE(64, 6, 1) NFLGPU_NTT_CASE(64, 7, 1) NFLGPU_NTT_CASE(64, 8, 1) NFLGPU_NTT_CASE(64, 9, 1)
    NFLGPU_NTT_CASE(64, 10, 1) NFLGPU_NTT_CASE(64, 11, 1) NFLGPU_NTT_CASE(64, 12, 1) NFLGPU_NTT_CASE(64, 13, 1)
    NFLGPU_NTT_CASE(64, 14, 1)
    NFLGPU_NTT_CASE(64, 15, 1) NFLGPU_NTT_CASE(64, 16, 1) NFLGPU_NTT_CASE(64, 17, 1) NFLGPU_NTT_CASE(64, 18, 1)
    NFLGPU_NTT_CASE(64, 19, 1) NFLGPU_NTT_CASE(64, 20, 1)
  }
  return cudaErrorInvalidValue;
}
}  // namespace nflgpu

// 64-bit limbs, fwd direction: degrees 2^2 .. 2^20 = params<uint64_t>::kMaxPolyDegree (above 2^14 the leading passes run as global-memory kernels, ntt_plan.h).
#include "ntt_launch.cuh"
namespace nflgpu {
cudaError_t launch_ntt_u64_fwd(int log2_degree, const NttLaunch &l, int device, int num_sms, cudaStream_t stream) {
  switch (log2_degree) {
    NFLGPU_NTT_CASE(64, 2, 0) NFLGPU_NTT_CASE(64, 3, 0) NFLGPU_NTT_CASE(64, 4, 0) NFLGPU_NTT_CASE(64, 5, 0)
    NFLGPU_NTT_CASE(64, 6, 0) NFLGPU_NTT_CASE(64, 7, 0) NFLGPU_NTT_CASE(64, 8, 0) NFLGPU_NTT_CASE(64, 9, 0)
    NFLGPU_NTT_CASE(64, 10, 0) NFLGPU_NTT_CASE(64, 11, 0) NFLGPU_NTT_CASE(64, 12, 0) NFLGPU_NTT_CASE(64, 13, 0)
    NFLGPU_NTT_CASE(64, 14, 0)
    NFLGPU_NTT_CASE(64, 15, 0) NFLGPU_NTT_CASE(64, 16, 0) NFLGPU_NTT_CASE(64, 17, 0) NFLGPU_NTT_CASE(64, 18, 0)
    NFLGPU_NTT_CASE(64, 19, 0) NFLGPU_NTT_CASE(64, 20, 0)
  }
  return cudaErrorInvalidValue;
}
}  // namespace nflgpu
#ifdef NFLGPU_TRACE
extern "C" int nflgpu_debug_trace(unsigned long long *out, int n, int reset) {
  if (out && cudaMemcpyFromSymbol(out, nflgpu::nflgpu_trace_buf, sizeof(unsigned long long) * (size_t)n) != cudaSuccess) return -3;
  if (reset) {
    static unsigned long long init[1 + 8192];
    init[0] = ~0ull;
    if (cudaMemcpyToSymbol(nflgpu::nflgpu_trace_buf, init, sizeof(init)) != cudaSuccess) return -3;
  }
  return 0;
}
#endif

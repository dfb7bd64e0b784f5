// Internal launcher interface of the NTT kernels (ntt_u{16,32,64}_{fwd,inv}.cu).
#ifndef NFLGPU_NTT_DISPATCH_H
#define NFLGPU_NTT_DISPATCH_H
#include <cstdint>
#include <cuda_runtime.h>

namespace nflgpu {

struct NttLaunch {
  const void *src;
  void *dst;
  const void *tw;      // device TW[nmoduli][N] of the requested direction
  const void *moduli;  // device Word[nmoduli]
  uint32_t nmoduli, batch;
};

// Returns cudaErrorInvalidValue when (limb_bits, log2_degree) has no kernel.
cudaError_t launch_ntt(int limb_bits, int log2_degree, bool inverse, const NttLaunch &l, int device, int num_sms,
                       cudaStream_t stream);
bool ntt_supported(int limb_bits, int log2_degree);

#define NFLGPU_DECL_LAUNCHER(name) \
  cudaError_t name(int log2_degree, const NttLaunch &l, int device, int num_sms, cudaStream_t stream);
NFLGPU_DECL_LAUNCHER(launch_ntt_u64_fwd) NFLGPU_DECL_LAUNCHER(launch_ntt_u64_inv)
NFLGPU_DECL_LAUNCHER(launch_ntt_u32_fwd) NFLGPU_DECL_LAUNCHER(launch_ntt_u32_inv)
NFLGPU_DECL_LAUNCHER(launch_ntt_u16_fwd) NFLGPU_DECL_LAUNCHER(launch_ntt_u16_inv)

}  // namespace nflgpu
#endif

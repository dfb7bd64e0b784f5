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
  const void *other = nullptr;        // fused epilogue operand (mode 2 only)
  const uint64_t *consts = nullptr;   // Barrett constants (mode 2 only)
  uint32_t *sched = nullptr;  // nmoduli + 1 zeroed counters for this launch (dynamic unit scheduling)
};

// Returns cudaErrorInvalidValue when (limb_bits, log2_degree) has no kernel.
// mode: 0 forward, 1 inverse, 2 forward fused with "* other"
cudaError_t launch_ntt(int limb_bits, int log2_degree, int mode, const NttLaunch &l, int device, int num_sms,
                       cudaStream_t stream);
bool ntt_supported(int limb_bits, int log2_degree);

#define NFLGPU_DECL_LAUNCHER(name) \
  cudaError_t name(int log2_degree, const NttLaunch &l, int device, int num_sms, cudaStream_t stream);
NFLGPU_DECL_LAUNCHER(launch_ntt_u64_fwd) NFLGPU_DECL_LAUNCHER(launch_ntt_u64_inv)
NFLGPU_DECL_LAUNCHER(launch_ntt_u32_fwd) NFLGPU_DECL_LAUNCHER(launch_ntt_u32_inv)
NFLGPU_DECL_LAUNCHER(launch_ntt_u16_fwd) NFLGPU_DECL_LAUNCHER(launch_ntt_u16_inv)
NFLGPU_DECL_LAUNCHER(launch_ntt_u64_fwdmul) NFLGPU_DECL_LAUNCHER(launch_ntt_u32_fwdmul) NFLGPU_DECL_LAUNCHER(launch_ntt_u16_fwdmul)

}  // namespace nflgpu
#endif

// Shared launcher body for one (limb, direction) translation unit.
#ifndef NFLGPU_NTT_LAUNCH_CUH
#define NFLGPU_NTT_LAUNCH_CUH
#include "ntt_dispatch.h"
#include "ntt_engine.cuh"
#include "ntt_cluster.cuh"
#include <cstdlib>

namespace nflgpu {

// launches the global-memory passes [FIRST, LAST] of a split transform, ascending (forward) or descending (inverse)
template <int LB, int LOGN, int PASS, bool INV> cudaError_t launch_gpasses(NttArgs a, int last, int num_sms, cudaStream_t stream) {
  typedef NttCfg<LB, LOGN> C;
  if constexpr (PASS >= 0 && PASS < C::SPLIT) {
    const uint64_t total = ((uint64_t)a.batch * a.nmoduli) << (C::n - C::e);
    uint64_t blocks = (total + 255) / 256;
    if (blocks > (uint64_t)num_sms * 16) blocks = (uint64_t)num_sms * 16;
    ntt_gpass_kernel<LB, LOGN, PASS, INV><<<(unsigned)blocks, 256, 0, stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || PASS == last) return e;
    a.src = a.dst;  // later passes work in place on dst
    return launch_gpasses<LB, LOGN, INV ? PASS - 1 : PASS + 1, INV>(a, last, num_sms, stream);
  } else {
    return cudaSuccess;
  }
}

// Cluster path for split transforms with one leading pass (ntt_cluster.cuh).  Returns true when the launch was made (or
// failed: *err); false when clusters of this shape cannot be scheduled on the device or NFLGPU_NO_CLUSTER=1 asks for the
// round-1 path (global-memory pass + tile kernel), which stays as the fallback and as the A/B partner.
template <int LB, int LOGN, int MODE> bool launch_ntt_cluster(const NttLaunch &l, int device, cudaStream_t stream, cudaError_t *err) {
  typedef ClusterCfg<LB, LOGN> C;
  if constexpr (!C::OK) {
    return false;
  } else {
    void (*kernel)(const ClusterArgs);
    if constexpr (MODE == 1) kernel = ntt_cluster_inv_kernel<LB, LOGN>;
    else if constexpr (MODE == 2) kernel = ntt_cluster_fwd_kernel<LB, LOGN, true>;
    else kernel = ntt_cluster_fwd_kernel<LB, LOGN, false>;
    static int max_clusters[64] = {0};  // per device: 0 = not asked yet, -1 = unusable
    if (device < 0 || device >= 64) return false;
    static const bool disabled = [] { const char *e = std::getenv("NFLGPU_NO_CLUSTER"); return e && e[0] == '1'; }();
    if (disabled) return false;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C::CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(C::THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = stream; cfg.attrs = attr; cfg.numAttrs = 1;
    if (max_clusters[device] == 0) {
      int n = 0;
      cfg.gridDim = dim3(C::CL);
      if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES) != cudaSuccess ||
          cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        max_clusters[device] = -1;
      } else {
        max_clusters[device] = n;
      }
    }
    if (max_clusters[device] < 0) return false;
    const uint64_t units = (uint64_t)l.batch * l.nmoduli;
    if (units == 0) { *err = cudaSuccess; return true; }
    if (units > 0xffffffffull) return false;
    const uint32_t ncl = (uint32_t)(units < (uint64_t)max_clusters[device] ? units : (uint64_t)max_clusters[device]);
    ClusterArgs a;
    a.src = l.src; a.dst = l.dst; a.tw = l.tw; a.moduli = l.moduli; a.nmoduli = l.nmoduli; a.batch = l.batch; a.nclusters = ncl;
    a.other = l.other; a.consts = l.consts;
    cfg.gridDim = dim3(ncl * C::CL);
    *err = cudaLaunchKernelEx(&cfg, kernel, a);
    return true;
  }
}

// MODE: 0 forward, 1 inverse, 2 forward with the fused "* other" epilogue
template <int LB, int LOGN, int MODE> cudaError_t launch_ntt_one(const NttLaunch &l, int device, int num_sms, cudaStream_t stream) {
  typedef NttCfg<LB, LOGN> C;
  constexpr size_t SMEM = MODE == 1 ? C::SMEM_BYTES_INV : C::SMEM_BYTES;  // (the inverse kernel may stage a larger twiddle table)
  void (*kernel)(const NttArgs);
  if constexpr (MODE == 1) kernel = ntt_inv_kernel<LB, LOGN>;
  else if constexpr (MODE == 2) kernel = ntt_fwd_kernel<LB, LOGN, true>;
  else kernel = ntt_fwd_kernel<LB, LOGN, false>;
  // per-device one-time setup: opt in to the shared-memory size and ask the occupancy calculator
  // (racing first calls from two host threads would both compute the same value; the store is a plain int)
  static int blocks_per_sm[64] = {0};
  if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
  if (blocks_per_sm[device] == 0) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, C::THREADS, SMEM);
    if (e != cudaSuccess) return e;
    blocks_per_sm[device] = occ > 0 ? occ : 1;
  }
  if (l.batch == 0) return cudaSuccess;
  if constexpr (C::SPLIT == 1) {  // 64-bit N = 2^15, 2^16: the whole unit on chip in a thread-block cluster
    cudaError_t ce = cudaSuccess;
    if (launch_ntt_cluster<LB, LOGN, MODE>(l, device, stream, &ce)) return ce;
  }
  // the kernels index sub-blocks with 32 bits: larger batches go out as several launches
  constexpr uint32_t kMaxBatch = (1u << 30) >> C::LOGG;
  if (l.batch > kMaxBatch) {
    const size_t poly_bytes = (size_t)l.nmoduli * C::N * sizeof(typename C::Store);
    for (uint32_t done = 0; done < l.batch; done += kMaxBatch) {
      NttLaunch part = l;
      part.batch = l.batch - done < kMaxBatch ? l.batch - done : kMaxBatch;
      part.src = static_cast<const char *>(l.src) + (size_t)done * poly_bytes;
      part.dst = static_cast<char *>(l.dst) + (size_t)done * poly_bytes;
      if (l.other) part.other = static_cast<const char *>(l.other) + (size_t)done * poly_bytes;
      cudaError_t e = launch_ntt_one<LB, LOGN, MODE>(part, device, num_sms, stream);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  }
  // persistent CTAs, each bound to one residue: grid = nmoduli * ctas_per_residue ~ one full wave
  const uint32_t resident = (uint32_t)num_sms * blocks_per_sm[device];
  uint32_t cpr = resident / l.nmoduli;
  if (cpr == 0) cpr = 1;
  const uint64_t need = ((((uint64_t)l.batch) << C::LOGG) + C::SLOTS - 1) / C::SLOTS;
  if (cpr > need) cpr = (uint32_t)need;
  NttArgs a;
  a.src = l.src; a.dst = l.dst; a.tw = l.tw; a.moduli = l.moduli;
  a.nmoduli = l.nmoduli; a.batch = l.batch; a.ctas_per_residue = cpr;
  a.other = l.other; a.consts = l.consts; a.sched = l.sched;
  if constexpr (C::SPLIT > 0 && MODE != 1) {  // forward: global passes 0 .. SPLIT-1 (src -> dst), then the tile kernel in place
    cudaError_t e = launch_gpasses<LB, LOGN, 0, false>(a, C::SPLIT - 1, num_sms, stream);
    if (e != cudaSuccess) return e;
    a.src = a.dst;
  }
  uint32_t grid = cpr * l.nmoduli;
  if constexpr (C::HOP && MODE != 1) {  // forward CTAs move from residue to residue: one full wave on every SM, whatever nmoduli is
    const uint64_t need_all = need * l.nmoduli;
    grid = need_all < resident ? (uint32_t)need_all : resident;
    if (grid == 0) grid = 1;
  }
  kernel<<<grid, C::THREADS, SMEM, stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if constexpr (C::SPLIT > 0 && MODE == 1) {  // inverse: tile kernel (src -> dst), then global passes SPLIT-1 .. 0 in place
    if (e != cudaSuccess) return e;
    a.src = a.dst;
    e = launch_gpasses<LB, LOGN, C::SPLIT - 1, true>(a, 0, num_sms, stream);
  }
  return e;
}

}  // namespace nflgpu

// NFLGPU_ONLY_LOGN restricts the instantiations to one size (fast experiment builds, tools/variants.sh)
#ifdef NFLGPU_ONLY_LOGN
#define NFLGPU_NTT_CASE(LB, LOGN, MODE) \
  case LOGN: if constexpr (LOGN == NFLGPU_ONLY_LOGN) return launch_ntt_one<LB, LOGN, MODE>(l, device, num_sms, stream); else break;
#else
#define NFLGPU_NTT_CASE(LB, LOGN, MODE) \
  case LOGN: return launch_ntt_one<LB, LOGN, MODE>(l, device, num_sms, stream);
#endif

#endif

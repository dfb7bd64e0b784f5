// Transforms larger than one shared-memory tile, on chip: a thread-block CLUSTER owns one (residue, polynomial) unit.
//
// 64-bit transforms of N = 2^15 (256 KiB per unit) and above do not fit the 227 KiB of one SM.  Round 1 ran their
// leading pass as a separate registers <-> HBM kernel (ntt_gpass_kernel): two full HBM round trips per transform.  Here a
// cluster of CL CTAs (sm_100a thread-block clusters, distributed shared memory; enabled for the CTA pair of N = 2^15, see
// ClusterCfg::OK) keeps the whole unit on chip:
//
//   forward   pass 0: every thread loads its E coefficients (stride N/E, lane-contiguous) straight from HBM, runs the pass's
//             butterflies in registers and SCATTERS the results into the sub-block tiles -- register k of a thread belongs to
//             sub-block k, which lives in the shared memory of CTA k / (sub-blocks per CTA): local stores for its own CTA,
//             st.shared::cluster (DSMEM) for the others;  cluster barrier;
//             passes 1..: each CTA finishes its own sub-blocks exactly like a small transform (FwdChain on a padded tile,
//             twiddles indexed by the sub-block number) and copies them out with coalesced 16-byte stores.
//   inverse   the mirror image: tiles <- HBM, passes NP-1 .. 1 in the tiles, cluster barrier, pass 0 GATHERS its E inputs from
//             the tiles of all CTAs (ld.shared::cluster), and writes its results to HBM.
//
// One read and one write of every coefficient per transform, as for the sizes that fit one tile.  Replaces the reference's
// core::ntt / inv_ntt (core.hpp:455-557) for the largest configuration of its own test matrix (tests/CMakeLists.txt:1-7:
// degree 32768, 124 bits, uint64_t).  A second cluster barrier per unit keeps the next unit's scatter (gather) away from tiles
// that are still being read (written).
#ifndef NFLGPU_NTT_CLUSTER_CUH
#define NFLGPU_NTT_CLUSTER_CUH

#include <cooperative_groups.h>
#include "ntt_engine.cuh"

namespace nflgpu {

namespace cg = cooperative_groups;

// Geometry of the cluster kernels for a split transform with exactly one leading pass (NttCfg::SPLIT == 1).
template <int LB, int LOGN> struct ClusterCfg : NttCfg<LB, LOGN> {
  typedef NttCfg<LB, LOGN> Base;
  typedef typename Base::Word Word;
  static constexpr int NSUB = 1 << Base::LOGG;                                                     // sub-blocks per unit
  static constexpr int CL = (int)(((size_t)Base::N * sizeof(Word) + 131071) / 131072);            // CTAs per cluster: 128 KiB of coefficients each
  static constexpr int TPC = NSUB / CL;                                                            // tiles (sub-blocks) per CTA
  static constexpr int SLOTS = TPC;                                                                // (unit_sync: named barrier per tile)
  static constexpr int THREADS = Base::TPU * TPC;                                                  // = N / (E * CL): also the pass-0 threads per CTA
  static constexpr bool DYNAMIC = false;
  static constexpr size_t SMEM_BYTES = (size_t)TPC * Base::TILE_WORDS * sizeof(Word);
  // Measured on B200 (profiles/r02_variants.log): N = 2^15 (2 CTAs x 512 threads, 32 coefficients per thread) 210 / 217 us against
  // 247 / 257 us for the global-memory pass + tile kernel (M = 2, batch 256); N = 2^16 (4 CTAs x 1024 threads, 16 coefficients) 253 / 267
  // against 217 / 221 -- four-CTA clusters with two barriers per unit lose, so only the CTA pair is enabled (-DNFLGPU_CLUSTER_MAX=4 to retry).
#ifndef NFLGPU_CLUSTER_MAX
#define NFLGPU_CLUSTER_MAX 2
#endif
  static constexpr bool OK = Base::SPLIT == 1 && NSUB % CL == 0 && CL <= NFLGPU_CLUSTER_MAX && THREADS <= 1024 && THREADS * CL == (Base::N >> Base::e) &&
                             SMEM_BYTES <= 227 * 1024 && Base::LOGG == Base::e;  // pass 0 runs all e stages: register k <-> sub-block k
};

struct ClusterArgs {
  const void *src;
  void *dst;
  const void *tw;
  const void *moduli;
  uint32_t nmoduli, batch, nclusters;
  const void *other;
  const uint64_t *consts;
};

// Distributed shared memory: a shared::cta address of this CTA mapped into the shared::cluster window of CTA `rank` (mapa), and
// loads / stores through that window.  Tile `sub` of the unit lives in CTA sub / TPC at local tile sub % TPC, so a thread needs
// one mapped base per CTA of the cluster (CL of them) and compile-time offsets for the tiles inside it.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, uint64_t v) { asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }
__device__ __forceinline__ void st_cluster(uint32_t addr, uint32_t v) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void ld_cluster(uint32_t addr, uint64_t &v) { asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory"); }
__device__ __forceinline__ void ld_cluster(uint32_t addr, uint32_t &v) { asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); }

template <int LB, int LOGN, bool MUL>
__global__ void __launch_bounds__(ClusterCfg<LB, LOGN>::THREADS, 1) ntt_cluster_fwd_kernel(const ClusterArgs a) {
  typedef ClusterCfg<LB, LOGN> C;
  typedef typename C::Word Word;
  typedef typename C::Store Store;
  typedef typename C::TW TW;
  extern __shared__ __align__(128) unsigned char smem[];
  Word *tiles = reinterpret_cast<Word *>(smem);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const uint32_t cid = blockIdx.x / C::CL;
  const int t0 = rank * C::THREADS + (int)threadIdx.x;          // pass 0: thread index inside the unit = position t0 of every sub-block
  const int slot = threadIdx.x / C::TPU, tl = threadIdx.x % C::TPU;
  const int lane_base = (threadIdx.x & 31) - (tl & 31);
  const int sub = rank * C::TPC + slot;                          // passes 1..: this thread's sub-block
  const int tid = sub * C::TPU + tl;
  Word *tile = tiles + (size_t)slot * C::TILE_WORDS;
  const Store *src = reinterpret_cast<const Store *>(a.src);
  Store *dst = reinterpret_cast<Store *>(a.dst);
  // scatter targets of pass 0: sub-block k, position t0 (the same offset in every tile): one mapped base per CTA of the cluster
  uint32_t cbase[C::CL];
#pragma unroll
  for (int r = 0; r < C::CL; ++r) cbase[r] = mapa_u32(smem_u32(tiles + C::pad(t0)), (uint32_t)r);

  const uint32_t units = a.batch * a.nmoduli;
  cluster.sync();  // every CTA of the cluster is running (its shared memory exists) before the first remote access
  for (uint32_t u = cid; u < units; u += a.nclusters) {
    const uint32_t cm = u % a.nmoduli;
    const size_t ubase = (size_t)u * C::N;
    const TW *tw = reinterpret_cast<const TW *>(a.tw) + (size_t)cm * C::N;
    const Word p = reinterpret_cast<const Word *>(a.moduli)[cm], twop = 2 * p, np = opaque_neg(p);
    Word x[C::E];
#pragma unroll
    for (int k = 0; k < C::E; ++k) x[k] = (Word)ld_coef(src + ubase + pass_pos<C, 0>(t0, k));
    fwd_pass<C, 0>(x, pass_tw<C, 0>(tw, t0), np, twop);
#pragma unroll
    for (int k = 0; k < C::E; ++k) st_cluster(cbase[k / C::TPC] + (uint32_t)((k % C::TPC) * C::TILE_WORDS * sizeof(Word)), x[k]);
    cluster.sync();  // every tile of the unit is complete (release / acquire across the cluster)
    FwdChain<C, 1>::run(x, tile, tw, p, np, twop, tid, slot, lane_base);
    unit_sync<C>(slot, lane_base);
    const size_t bbase = ubase + (size_t)sub * C::B;
    if (MUL) tile_to_gmem_mul<C, LB>(tile, dst + bbase, reinterpret_cast<const Store *>(a.other) + bbase, tl, p, a.consts[cm]);
    else tile_to_gmem<C>(tile, dst + bbase, tl);
    cluster.sync();  // all tiles have been copied out before the next unit scatters into them
  }
}

template <int LB, int LOGN>
__global__ void __launch_bounds__(ClusterCfg<LB, LOGN>::THREADS, 1) ntt_cluster_inv_kernel(const ClusterArgs a) {
  typedef ClusterCfg<LB, LOGN> C;
  typedef typename C::Word Word;
  typedef typename C::Store Store;
  typedef typename C::TW TW;
  extern __shared__ __align__(128) unsigned char smem[];
  Word *tiles = reinterpret_cast<Word *>(smem);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const uint32_t cid = blockIdx.x / C::CL;
  const int t0 = rank * C::THREADS + (int)threadIdx.x;
  const int slot = threadIdx.x / C::TPU, tl = threadIdx.x % C::TPU;
  const int lane_base = (threadIdx.x & 31) - (tl & 31);
  const int sub = rank * C::TPC + slot;
  const int tid = sub * C::TPU + tl;
  Word *tile = tiles + (size_t)slot * C::TILE_WORDS;
  const Store *src = reinterpret_cast<const Store *>(a.src);
  Store *dst = reinterpret_cast<Store *>(a.dst);
  uint32_t cbase[C::CL];
#pragma unroll
  for (int r = 0; r < C::CL; ++r) cbase[r] = mapa_u32(smem_u32(tiles + C::pad(t0)), (uint32_t)r);

  const uint32_t units = a.batch * a.nmoduli;
  // Experiment (-DNFLGPU_PIPE=1), pipelining across units like ntt_inv_kernel (NttCfg::PIPE_INV): this CTA's sub-blocks of the NEXT unit are
  // copied into the tiles with cp.async as soon as every CTA of the cluster has gathered its pass-0 window out of them, and land while
  // pass 0 computes and stores.  Measured on B200 (profiles/r02_variants.log block 11): 217.4 vs 216.6 us (M = 2, batch 256), 833.6 vs
  // 824.8 us (M = 4, batch 512) -- no gain, the cluster barrier moved in front of pass 0 costs what the hidden copy saves; off by default.
#if defined(NFLGPU_PIPE) && NFLGPU_PIPE == 1
  constexpr bool PIPE = sizeof(Store) == sizeof(Word);
#else
  constexpr bool PIPE = false;
#endif
  cluster.sync();  // every CTA of the cluster is running (its shared memory exists) before the first remote access
  if constexpr (PIPE) {
    if (cid < units) gmem_to_tile_async<C>(tile, src + (size_t)cid * C::N + (size_t)sub * C::B, tl);
  }
  for (uint32_t u = cid; u < units; u += a.nclusters) {
    const uint32_t cm = u % a.nmoduli;
    const size_t ubase = (size_t)u * C::N;
    const TW *tw = reinterpret_cast<const TW *>(a.tw) + (size_t)cm * C::INV_TW;
    const TW ninv = __ldg(tw + C::N - 1);
    const Word p = reinterpret_cast<const Word *>(a.moduli)[cm], twop = 2 * p, np = opaque_neg(p);
    Word x[C::E];
    if constexpr (PIPE) cp_async_wait();
    else gmem_to_tile<C>(tile, src + ubase + (size_t)sub * C::B, tl);
    unit_sync<C>(slot, lane_base);
    InvChain<C, C::NP - 1>::run(x, tile, tw, p, np, twop, ninv, tid, slot, lane_base);  // passes NP-1 .. 2, tile -> tile
    unit_sync<C>(slot, lane_base);
    tile_load<C, 1>(x, tile, tid);                                                       // pass 1, back into the tile
    inv_pass<C, 1>(x, pass_tw<C, 1>(tw, tid), p, np, twop, ninv);
    tile_store<C, 1>(x, tile, tid);
    cluster.sync();  // every sub-block of the unit has finished its tile passes
#pragma unroll
    for (int k = 0; k < C::E; ++k) ld_cluster(cbase[k / C::TPC] + (uint32_t)((k % C::TPC) * C::TILE_WORDS * sizeof(Word)), x[k]);
    if constexpr (PIPE) {
      cluster.sync();  // all gathers are done: the tiles are free for the next unit
      const uint32_t un = u + a.nclusters;
      if (un < units) gmem_to_tile_async<C>(tile, src + (size_t)un * C::N + (size_t)sub * C::B, tl);
    }
    inv_pass<C, 0>(x, pass_tw<C, 0>(tw, t0), p, np, twop, ninv);
#pragma unroll
    for (int k = 0; k < C::E; ++k) dst[ubase + pass_pos<C, 0>(t0, k)] = (Store)x[k];
    if constexpr (!PIPE) cluster.sync();  // all gathers are done before the next unit overwrites the tiles
  }
}

}  // namespace nflgpu
#endif

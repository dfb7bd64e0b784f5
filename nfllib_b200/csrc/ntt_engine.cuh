// Multi-pass register-radix negacyclic NTT kernels for sm_100a (integer pipes + shared memory; no tensor cores:
// a butterfly network is not a dense contraction).
//
// Replaces, for device-resident batches, the reference's
//   poly::core::ntt_pow_phi      core.hpp:594-600  (phi twist, ops.hpp:227-242  +  core::ntt, core.hpp:455-532,
//                                                    ntt_loop / ntt_loop_body algos.hpp:16-73, sse.hpp:162-301,
//                                                    avx2.hpp:160-302)
//   poly::core::invntt_pow_invphi core.hpp:608-614  (inv_ntt core.hpp:539-557, permut.hpp:13-117, N^-1 phi^-i twist)
// with the merged-psi formulation (tables.cpp): forward = Cooley-Tukey butterflies, natural order in, the
// reference's bit-reversed evaluation order out; inverse = Gentleman-Sande butterflies consuming that order.
// No twist pass, no bit-reversal pass, no separate correction pass: one read and one write of every
// coefficient per transform.
//
// Work decomposition: one (residue, polynomial) unit = a contiguous slab of N limbs.  A CTA is bound to one
// residue (its twiddles are staged once into shared memory by a TMA bulk copy when they fit), and holds
// SLOTS groups of TPU = N/E threads; each group walks over the polynomials of the batch.  Inside a unit each
// thread keeps E = 2^e coefficients in registers and runs up to e butterfly stages per pass (ntt_plan.h);
// passes exchange coefficients through a padded shared-memory tile (rows of E words + 16 bytes, so both the
// row-per-thread and the lane-contiguous access patterns are bank-conflict free).
#ifndef NFLGPU_NTT_ENGINE_CUH
#define NFLGPU_NTT_ENGINE_CUH

#include "modarith.cuh"
#include "modmul.cuh"
#include "ntt_plan.h"

namespace nflgpu {

struct NttArgs {
  const void *src;
  void *dst;
  const void *tw;      // TW[nmoduli][N] for this direction
  const void *moduli;  // Word[nmoduli]
  uint32_t nmoduli;
  uint32_t batch;
  uint32_t ctas_per_residue;
  // fused epilogue of the forward kernel (nflgpu_polymul): dst = ntt_pow_phi(src) * other, coefficient-wise
  const void *other;       // Store[batch][nmoduli][N], canonical, NTT domain; null when unused
  const uint64_t *consts;  // Barrett constants per residue (pointwise.h)
  // dynamic unit scheduling (NttCfg::DYNAMIC): sched[cm] = next unclaimed sub-block of residue cm, sched[nmoduli] = CTAs done;
  // all zero between launches (the last CTA to finish resets them)
  uint32_t *sched;
};

template <int LB, int LOGN> struct NttCfg {
  typedef Arith<LB> A;
  typedef typename A::Word Word;
  typedef typename A::Store Store;
  typedef typename A::TW TW;
  static constexpr int WB = A::WORD_BITS;
  static constexpr int n = LOGN;
  static constexpr int N = 1 << n;
  static constexpr int e = plan_e(n, WB);
  static constexpr int E = 1 << e;
  // Forward butterflies of the 16-coefficient 64-bit kernels use the "top-bit" lazy range (modarith.cuh csub_top): values
  // anywhere in [0, 2^64), conditional subtract in four instructions instead of five.  Measured on B200
  // (profiles/r01e_variants.log): N = 1024: 126.8 -> 124.1 us per launch; the 32-coefficient kernels (4 warps per
  // sub-partition) lose 10-13 % with it -- ptxas routes every predicated carry through one predicate register, which
  // serialises the sixteen conditional subtracts of a stage -- so they keep the [0, 4p) form.
#if defined(NFLGPU_LAZY64)
  static constexpr bool TOP = NFLGPU_LAZY64 != 0 && WB == 64;                       // experiment builds: 0 off, 1 everywhere,
  static constexpr bool TOP_SELECT = NFLGPU_LAZY64 == 2 && WB == 64 && e > 4;         // 2 = select form on the 32-coefficient kernels
#else
  static constexpr bool TOP = WB == 64 && e <= 4;
  static constexpr bool TOP_SELECT = false;
#endif
  static constexpr int NP = plan_npass(n, WB);
  static constexpr int SPLIT = plan_split(n, WB);        // leading passes run as global-memory kernels (0 unless N is huge)
  static constexpr int LOGB = plan_hi(n, WB, SPLIT);     // the tile kernel transforms sub-blocks of B = 2^LOGB words
  static constexpr int B = 1 << LOGB;
  static constexpr int LOGG = n - LOGB;                  // sub-blocks per unit = 2^LOGG
  static constexpr int TPU = B >> e;                     // threads per sub-block (per unit when SPLIT == 0)
  static constexpr int VEC = 16 / (int)sizeof(Word);    // words per 16-byte shared-memory vector
  static constexpr int PADW = VEC;                       // 16 bytes of padding per row of E words
  static constexpr int ROW = E + PADW;
  // 32-bit words, N = 4096, shape (4,4,4): rows of 16 words + 16 bytes make the lane-contiguous accesses of passes 0 and 1 and the
  // 16-byte copies two wavefronts instead of one (tools/bank_conflicts.py; ncu counted 16 M conflicts per launch of the C4 shape).
  // This shape uses an unpadded tile with an XOR swizzle instead: address bits 2,3 ^= position bits 5,6 and bit 4 ^= bit 8, which
  // is conflict free for every pass and keeps 16-byte vectors contiguous (swz / swz_k below).
#ifdef NFLGPU_NO_SWIZZLE
  static constexpr bool SWZ = false;
#else
  static constexpr bool SWZ = WB == 32 && n == 12 && e == 4 && SPLIT == 0;
#endif
  static constexpr int TILE_WORDS = (NP - SPLIT > 1) ? (SWZ ? B : (B >> e) * ROW) : 0;
  // CTA size: 256 threads (2+ CTAs per SM)
#ifdef NFLGPU_TARGET_THREADS
  static constexpr int TARGET_THREADS = NFLGPU_TARGET_THREADS;
#else
  // N = 1024 x 64-bit: ONE CTA per SM.  Measured (profiles/r02_variants.log): 2 x 512 threads 120.3 / 121.9 us (forward / inverse),
  // 1 x 1024 threads (64 registers) 115.6 / 117.2, 1 x 896 threads = 14 two-warp units with 72 registers 114.1 / 114.4, 1 x 768 (85
  // registers) 125.4 / 121.4: the kernel is instruction-issue bound, so seven warps per sub-partition with 2 % fewer instructions win
  static constexpr int TARGET_THREADS = (LB == 64 && LOGN == 10) ? 896 : 256;
#endif
  static constexpr int SLOTS = (TPU >= TARGET_THREADS) ? 1 : TARGET_THREADS / TPU;
  static constexpr int THREADS = TPU * SLOTS;
#ifdef NFLGPU_MIN_BLOCKS
  static constexpr int MIN_BLOCKS = NFLGPU_MIN_BLOCKS;
#else
  static constexpr int MIN_BLOCKS = (THREADS <= 256) ? ((WB == 32 && E <= 32) ? 4 : 2) : 1;
#endif
  static constexpr bool TW_SMEM = SPLIT == 0 && (size_t)N * sizeof(TW) <= 32768;
  static constexpr size_t TW_BYTES = TW_SMEM ? (size_t)N * sizeof(TW) : 0;
  // Unit slots claim their next sub-block from a per-residue atomic counter instead of striding over the batch: a slot
  // that runs ahead (the warp schedulers are not fair) takes more units, so all resident warps stay busy until the batch
  // is exhausted (static striding left the average warp idle for the last ~20 % of the launch, ncu sm__warps_active).
#ifdef NFLGPU_DYNAMIC
  static constexpr bool DYNAMIC = NFLGPU_DYNAMIC != 0;
  static constexpr bool CLAIM_LATE = NFLGPU_DYNAMIC == 2;  // claim at the bottom of the iteration (no register carried through the unit)
#else
  // measured on B200 (profiles/r01d_kbench_all.txt): -4 .. -11 % time for N >= 2048 and for the 32-bit N = 1024 kernels, +3 % for
  // N = 1024 x 64-bit (its ~7 small units per slot already balance; the extra loop state costs instructions there)
  static constexpr bool DYNAMIC = LOGN >= 11 || (LB != 64 && LOGN == 10);
  static constexpr bool CLAIM_LATE = false;
#endif
  // A CTA whose twiddles come from global memory (tables too large for shared memory) is not tied to its residue: when the
  // units of its own residue are exhausted it moves on to the next one.  The grid can then use EVERY SM whatever nmoduli is
  // (8 moduli used to leave 148 - 8 * 18 = 4 SMs idle), and residues finish together.  Forward kernels only: measured on B200
  // (profiles/r02_variants.log) C3 forward 1426.8 -> 1407.8 us; the inverse kernels spill with the extra loop (C5 inverse 915.9 -> 983.6 us).
#ifdef NFLGPU_NO_HOP
  static constexpr bool HOP = false;
#else
  static constexpr bool HOP = DYNAMIC && !TW_SMEM;
#endif
  // Software pipelining across units: the inverse kernel copies the NEXT unit's slab into the tile with cp.async (no registers) while
  // the current unit's final pass computes and stores -- the tile is free from the moment every thread has loaded its last-pass
  // window -- and the forward kernel issues the next unit's pass-0 loads before the current unit's copy-out instead of after it.
  // Measured on B200 (profiles/r02_variants.log block 8, -DNFLGPU_PIPE=1 against the tree): inverse N = 4096 u64 -3.5 %, C4 (u32
  // N = 4096) -2.1 %, C5 -0.6 %, C3 -0.9 %, but N = 1024 u64 +1.5 % (its 14 units per SM already overlap); forward C4 -1.8 %, every
  // 64-bit size +1.1 .. +1.6 % (64 more live registers through the copy-out).  So: inverse from N = 4096 up, forward for C4's shape only.
  // -DNFLGPU_PIPE=0/1 forces it off / on everywhere.
#ifdef NFLGPU_PIPE
  static constexpr bool PIPE_INV = NFLGPU_PIPE != 0 && (NP - SPLIT) > 1 && sizeof(Store) == sizeof(Word);
  static constexpr bool PIPE_FWD = NFLGPU_PIPE != 0 && (NP - SPLIT) > 1;
#else
  static constexpr bool PIPE_INV = n >= 12 && (NP - SPLIT) > 1 && sizeof(Store) == sizeof(Word);
  static constexpr bool PIPE_FWD = WB == 32 && n == 12 && (NP - SPLIT) > 1;
#endif
  // Short first pass (r0 < e stages): the low e - r0 bits of the register index are independent columns.  With adjacent columns --
  // position = (k_hi << (n - r0)) | (tid << CB) | k_lo -- pass 0 moves them as 16-byte vectors: LDG.128 from HBM / STS.128 into the
  // tile in the forward kernel, LDS.128 / STG.128 in the inverse kernel, half the load / store instructions of the 8-byte
  // column-strided form (pass_pos, window_load / window_store below).  Measured on B200 against the strided form (-DNFLGPU_ADJ=0,
  // profiles/r02_variants.log block 9): N = 1024 (two column bits: 32 bytes per lane, two half-used sectors per request) +1.7 %
  // forward / +3.7 % inverse, N = 2048 +0.7 / -0.3 %, N = 16384 (one column bit, 16 bytes per lane) -0.4 / -1.4 %.  So it is on for
  // the N = 16384 shape only; -DNFLGPU_ADJ=1 turns it on wherever the tile layout stays conflict free (one column bit, or two with
  // 16-coefficient rows: tools/bank_conflicts.py), -DNFLGPU_ADJ=0 off everywhere.
  static constexpr int CB0 = (WB == 64 && SPLIT == 0 && NP > 1) ? e - plan_r(n, WB, 0) : 0;
#if defined(NFLGPU_ADJ) && NFLGPU_ADJ == 0
  static constexpr int CB = 0;
#elif defined(NFLGPU_ADJ)
  static constexpr int CB = (CB0 == 1 || (CB0 == 2 && e == 4)) ? CB0 : 0;
#else
  static constexpr int CB = (CB0 == 1 && n == 14) ? 1 : 0;
#endif
  static constexpr bool ADJ = CB >= 1;
  static constexpr size_t SCHED_BYTES = DYNAMIC ? (((size_t)2 * SLOTS * sizeof(uint32_t) + 15) & ~(size_t)15) : 0;
  // Experiment (-DNFLGPU_TMA_SLAB, north_star's "coefficients staged into shared memory via TMA"; statically walked single-tile shapes):
  // one elected thread per unit slot brings the next unit's slab into the slot's tile with a bulk copy (cp.async.bulk + mbarrier) as
  // soon as the copy-out has released the tile, and pass 0 reads its window from shared memory instead of global memory.  Measured on
  // B200 for N = 1024 x 64-bit: profiles/r02_variants.log block 10.  Not in the shipped build.
#ifdef NFLGPU_TMA_SLAB
  static constexpr bool TMA_SLAB = SPLIT == 0 && !DYNAMIC && NP > 1 && sizeof(Store) == sizeof(Word) && !ADJ;
#else
  static constexpr bool TMA_SLAB = false;
#endif
  static constexpr size_t SLAB_BAR_BYTES = TMA_SLAB ? (((size_t)SLOTS * 8 + 15) & ~(size_t)15) : 0;
  static constexpr size_t TILE_OFF = TW_BYTES + 16 /* mbarrier */ + SCHED_BYTES + SLAB_BAR_BYTES;
  static constexpr size_t SMEM_BYTES = TILE_OFF + (size_t)SLOTS * TILE_WORDS * sizeof(Word);
  // Inverse direction: N^-1 folded into the twiddles (ntt_plan.h plan_fold): N + N/2 table entries per residue, all of them staged
  // into shared memory when the table is staged at all.
  static constexpr bool FOLD = plan_fold(n, WB);
  static constexpr int INV_TW = plan_inv_entries(n, WB);
  static constexpr size_t TW_BYTES_INV = TW_SMEM ? (size_t)INV_TW * sizeof(TW) : 0;
  static constexpr size_t TILE_OFF_INV = TW_BYTES_INV + 16 /* mbarrier */ + SCHED_BYTES;
  static constexpr size_t SMEM_BYTES_INV = TILE_OFF_INV + (size_t)SLOTS * TILE_WORDS * sizeof(Word);
  // tile address of a position (only its offset inside the sub-block matters)
  static NFLGPU_DEVFN int pad(int pos) { return (pos & (B - 1)) + ((pos & (B - 1)) >> e) * PADW; }
  // the same for a compile-time position offset made of register-index bits only (below B by construction)
  static __host__ __device__ constexpr int pad_k(int off) { return off + (off >> e) * PADW; }
  // XOR-swizzled tile (SWZ): the swizzle term of a position, and the tile address of a position
  static __host__ __device__ constexpr int swz_x(int pos) { return (((pos >> 5) & 3) << 2) | (((pos >> 8) & 1) << 4); }
  static NFLGPU_DEVFN int swz(int pos) { return (pos & (B - 1)) ^ swz_x(pos); }
  // tile address of any position, whichever layout the configuration uses
  static NFLGPU_DEVFN int taddr(int pos) { return SWZ ? swz(pos) : pad(pos); }
};

// ---- small PTX helpers ---------------------------------------------------------------------------------

// Coefficient loads.  The kernels support dst == src (nflgpu_polymul's inverse, batch::ntt_pow_phi): every word of a slab is read
// exactly once, by the unit that later overwrites it, before any of that unit's stores -- the invariant that makes the
// non-coherent path (ld.global.nc) legal there.  -DNFLGPU_PLAIN_LD switches to ordinary loads (measured: no difference).
template <class T> NFLGPU_DEVFN T ld_coef(const T *p) {
#if defined(NFLGPU_PLAIN_LD) || !defined(__CUDA_ARCH__)
  return *p;
#else
  return __ldg(p);
#endif
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, void *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// L2 prefetch of the slab the slot will work on next (static unit walk only: the next index is known a whole unit ahead).
// One 128-byte line per thread and request; the later LDGs then come from L2 instead of HBM.
template <class C> __device__ __forceinline__ void prefetch_unit(const typename C::Store *slab, int tl) {
#if !defined(NFLGPU_PREFETCH) || NFLGPU_PREFETCH
  constexpr int LINES = (int)(C::B * sizeof(typename C::Store) / 128);
#pragma unroll
  for (int j = 0; j < (LINES + C::TPU - 1) / C::TPU; ++j) {
    const int line = tl + j * C::TPU;
    if (line < LINES) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(slab) + (size_t)line * 128));
  }
#endif
}

template <class C> __device__ __forceinline__ void unit_sync(int slot, int lane_base) {
#if defined(NFLGPU_ABL) && (NFLGPU_ABL & 4)  // timing experiment: no barriers between the passes (results are wrong)
  return;
#endif
  if (C::TPU >= 64) {
    if (C::SLOTS == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "n"(C::TPU) : "memory");
  } else if (C::TPU == 32) {
    __syncwarp();
  } else {
    __syncwarp((C::TPU >= 32 ? 0xffffffffu : ((1u << (C::TPU & 31)) - 1u)) << lane_base);
  }
}

// ---- butterfly networks on the register window -------------------------------------------------------------

// position (inside the unit) of register k of thread `tid` in pass PASS; `tid` is the thread's index inside the whole
// unit: for split transforms (sub-block number << log2(TPU)) | index inside the sub-block
template <class C, int PASS> NFLGPU_DEVFN int pass_pos(int tid, int k) {
  constexpr int hi = plan_hi(C::n, C::WB, PASS), c = plan_c(C::n, C::WB, PASS);
  if (PASS == 0 && C::ADJ) {  // adjacent columns: the butterfly bits of k on top, the thread in the middle, the column below
    constexpr int r0 = plan_r(C::n, C::WB, 0);
    return ((k >> C::CB) << (C::n - r0)) | (tid << C::CB) | (k & ((1 << C::CB) - 1));
  }
  const int g = tid >> c, l = tid & ((1 << c) - 1);
  return (g << hi) | (k << c) | l;
}

// Forward pass PASS: stages s0 .. s0+r-1, Cooley-Tukey, values lazily kept in [0, 4p) (32-bit words) or anywhere in
// [0, 2^64) (64-bit words, C::TOP).
// tw points at this pass's entry [0][g]; consecutive e_idx are G entries apart.
template <class C, int PASS> NFLGPU_DEVFN void fwd_pass(typename C::Word (&x)[C::E], const typename C::TW *tw,
                                                                      typename C::Word np, typename C::Word twop) {
  typedef typename C::A A;
  typedef typename C::Word Word;
  constexpr int r = plan_r(C::n, C::WB, PASS), G = 1 << plan_s0(C::n, C::WB, PASS), e = C::e;
  const Word n2p = np + np;  // -2p
#pragma unroll
  for (int q = 0; q < r; ++q) {
    const int bit = e - 1 - q;
#pragma unroll
    for (int k = 0; k < C::E; ++k) {
      if (k & (1 << bit)) continue;
      const int eidx = (1 << q) - 1 + (k >> (e - q));
      const typename C::TW t = tw[eidx * G];
      Word X = x[k];
      if (!(PASS == 0 && q == 0)) X = C::TOP ? (C::TOP_SELECT ? csub_top_select(X, n2p) : csub_top(X, n2p)) : csub_lazy(X, twop);  // first stage sees canonical input
      const Word T = A::mul_shoup_lazy(x[k | (1 << bit)], A::tw_w(t), A::tw_ws(t), np);
      x[k] = X + T;
      x[k | (1 << bit)] = C::TOP ? subadd(X, T, twop) : X - T + twop;
    }
  }
}

// Inverse pass PASS: stages s0+r-1 .. s0 (reverse order), Gentleman-Sande, values lazily kept in [0, 2p);
// the very last stage (PASS 0, q 0) produces canonical values.  N^-1: with folded tables (C::FOLD, ntt_plan.h plan_fold) the first
// pass to run (NP-1: the thread's window is the position's low e bits) takes the twiddle of every butterfly whose register-index
// bits below the paired one are all zero from the N^-1-scaled copy -- a compile-time choice -- and multiplies x[0] by N^-1 at its
// end, after which every coefficient of the unit carries the factor; otherwise the last stage multiplies every sum output by N^-1
// (its difference twiddle is pre-scaled in the table).
// (A top-bit variant of this direction -- sums held with the bias 2^63 - 2p so that "U + V >= 2p" is a sign bit -- was built
// and measured on B200: same instruction count gain as the forward one, no time gain at any size; not kept.)
template <class C, int PASS> NFLGPU_DEVFN void inv_pass(typename C::Word (&x)[C::E], const typename C::TW *tw,
                                                       typename C::Word p, typename C::Word np, typename C::Word twop,
                                                       const typename C::TW ninv) {
  typedef typename C::A A;
  typedef typename C::Word Word;
  constexpr int r = plan_r(C::n, C::WB, PASS), G = 1 << plan_s0(C::n, C::WB, PASS), e = C::e;
  constexpr bool FIRST = C::FOLD && PASS == C::NP - 1;
  const typename C::TW *twa = tw + (C::N - plan_off(C::n, C::WB, PASS));  // (FIRST only) the scaled copy of this pass's entries
#pragma unroll
  for (int q = r - 1; q >= 0; --q) {
    const int bit = e - 1 - q;
#pragma unroll
    for (int k = 0; k < C::E; ++k) {
      if (k & (1 << bit)) continue;
      const int eidx = (1 << q) - 1 + (k >> (e - q));
      // (the stage pairing bit 0 has the scaled values in the table proper: every butterfly of it qualifies)
      const bool scaled = FIRST && bit != 0 && (k & ((1 << bit) - 1)) == 0;
      const typename C::TW t = (scaled ? twa : tw)[eidx * G];
      const Word U = x[k], V = x[k | (1 << bit)];
      const Word D = A::mul_shoup_lazy(U - V + twop, A::tw_w(t), A::tw_ws(t), np);
      if (PASS == 0 && q == 0) {
        if (C::FOLD) x[k] = csub_lazy(csub_lazy(U + V, twop), p);
        else x[k] = csub_lazy(A::mul_shoup_lazy(U + V, A::tw_w(ninv), A::tw_ws(ninv), np), p);
        x[k | (1 << bit)] = csub_lazy(D, p);
      } else {
        x[k] = csub_lazy(U + V, twop);
        x[k | (1 << bit)] = D;
      }
    }
  }
  if (FIRST) {  // the one coefficient of the window no difference branch has touched
    x[0] = A::mul_shoup_lazy(x[0], A::tw_w(ninv), A::tw_ws(ninv), np);
    if (PASS == 0) x[0] = csub_lazy(x[0], p);  // (one-pass transforms: this is also the last stage)
  }
}

// lazy forward value -> canonical (core.hpp:523-529)
template <class C> NFLGPU_DEVFN typename C::Word fwd_canon(typename C::Word v, typename C::Word p, typename C::Word twop) {
  if (C::TOP) return canon_full(v, p, ((typename C::Word)1 << (C::WB - 2)) - p);
  return csub_lazy(csub_lazy(v, twop), p);
}

template <class C, int PASS> NFLGPU_DEVFN const typename C::TW *pass_tw(const typename C::TW *tw, int tid) {
  constexpr int c = plan_c(C::n, C::WB, PASS), off = plan_off(C::n, C::WB, PASS);
  return tw + off + (tid >> c);
}

// tile <-> registers for pass PASS.  The last pass (c == 0) owns a whole padded row: 16-byte vector accesses.
template <class C, int PASS> NFLGPU_DEVFN void tile_load(typename C::Word (&x)[C::E], const typename C::Word *tile, int tid) {
  typedef typename C::Word Word;
  constexpr int c = plan_c(C::n, C::WB, PASS);
  if (PASS == 0 && C::ADJ) {  // adjacent columns: 2^CB consecutive words per butterfly index, 16-byte vectors (padded layout only)
    constexpr int r0 = plan_r(C::n, C::WB, 0), COLS = 1 << C::CB;
    const Word *base = tile + C::pad(tid << C::CB);
#pragma unroll
    for (int kh = 0; kh < (1 << r0); ++kh) {
#pragma unroll
      for (int v = 0; v < COLS / C::VEC; ++v) {
        const uint4 t = *reinterpret_cast<const uint4 *>(base + C::pad_k(kh << (C::n - r0)) + v * C::VEC);
        const Word *w = reinterpret_cast<const Word *>(&t);
#pragma unroll
        for (int j = 0; j < C::VEC; ++j) x[kh * COLS + v * C::VEC + j] = w[j];
      }
    }
  } else if (C::SWZ) {
    // swz is linear over XOR: swz(T | K) = swz(T) ^ (K ^ swz_x(K)) for the thread part T and the register part K = k << c.
    // The bits of K outside [4:2] are disjoint from everything else (they add), the rest selects one of a few XOR variants of the
    // thread's base address: 2 in pass 0, 8 in pass 1, 4 (one per 16-byte vector) in pass 2.
    const int base = C::swz(pass_pos<C, PASS>(tid, 0));
    if (c == 0) {
#pragma unroll
      for (int v = 0; v < C::E / C::VEC; ++v) {
        const uint4 t = *reinterpret_cast<const uint4 *>(tile + (base ^ (v * C::VEC)));
        const Word *w = reinterpret_cast<const Word *>(&t);
#pragma unroll
        for (int j = 0; j < C::VEC; ++j) x[v * C::VEC + j] = w[j];
      }
    } else {
#pragma unroll
      for (int k = 0; k < C::E; ++k) x[k] = tile[(base ^ (((k << c) & 0x1c) ^ C::swz_x(k << c))) + ((k << c) & ~0x1c)];
    }
  } else if (c == 0) {
    const uint4 *row = reinterpret_cast<const uint4 *>(tile + (tid & (C::TPU - 1)) * C::ROW);
#pragma unroll
    for (int v = 0; v < C::E / C::VEC; ++v) {
      uint4 t = row[v];
      const Word *w = reinterpret_cast<const Word *>(&t);
#pragma unroll
      for (int j = 0; j < C::VEC; ++j) x[v * C::VEC + j] = w[j];
    }
  } else {
    // position bits of k and of the thread are disjoint, so pad(pos(tid, k)) = pad(pos(tid, 0)) + a compile-time offset:
    // one base register per pass and immediate offsets instead of E computed addresses
    const Word *base = tile + C::pad(pass_pos<C, PASS>(tid, 0));
#pragma unroll
    for (int k = 0; k < C::E; ++k) x[k] = base[C::pad_k(k << c)];
  }
}
template <class C, int PASS> NFLGPU_DEVFN void tile_store(const typename C::Word (&x)[C::E], typename C::Word *tile, int tid) {
  typedef typename C::Word Word;
  constexpr int c = plan_c(C::n, C::WB, PASS);
  if (PASS == 0 && C::ADJ) {  // (addresses as in tile_load)
    constexpr int r0 = plan_r(C::n, C::WB, 0), COLS = 1 << C::CB;
    Word *base = tile + C::pad(tid << C::CB);
#pragma unroll
    for (int kh = 0; kh < (1 << r0); ++kh) {
#pragma unroll
      for (int v = 0; v < COLS / C::VEC; ++v) {
        uint4 t;
        Word *w = reinterpret_cast<Word *>(&t);
#pragma unroll
        for (int j = 0; j < C::VEC; ++j) w[j] = x[kh * COLS + v * C::VEC + j];
        *reinterpret_cast<uint4 *>(base + C::pad_k(kh << (C::n - r0)) + v * C::VEC) = t;
      }
    }
  } else if (C::SWZ) {  // (addresses as in tile_load)
    const int base = C::swz(pass_pos<C, PASS>(tid, 0));
    if (c == 0) {
#pragma unroll
      for (int v = 0; v < C::E / C::VEC; ++v) {
        uint4 t;
        Word *w = reinterpret_cast<Word *>(&t);
#pragma unroll
        for (int j = 0; j < C::VEC; ++j) w[j] = x[v * C::VEC + j];
        *reinterpret_cast<uint4 *>(tile + (base ^ (v * C::VEC))) = t;
      }
    } else {
#pragma unroll
      for (int k = 0; k < C::E; ++k) tile[(base ^ (((k << c) & 0x1c) ^ C::swz_x(k << c))) + ((k << c) & ~0x1c)] = x[k];
    }
  } else if (c == 0) {
    uint4 *row = reinterpret_cast<uint4 *>(tile + (tid & (C::TPU - 1)) * C::ROW);
#pragma unroll
    for (int v = 0; v < C::E / C::VEC; ++v) {
      uint4 t;
      Word *w = reinterpret_cast<Word *>(&t);
#pragma unroll
      for (int j = 0; j < C::VEC; ++j) w[j] = x[v * C::VEC + j];
      row[v] = t;
    }
  } else {
    Word *base = tile + C::pad(pass_pos<C, PASS>(tid, 0));
#pragma unroll
    for (int k = 0; k < C::E; ++k) base[C::pad_k(k << c)] = x[k];
  }
}

// coalesced 16-byte-per-lane copies between the padded tile and the unit's slab in global memory
template <class C> __device__ __forceinline__ void tile_to_gmem(const typename C::Word *tile, typename C::Store *g, int tid) {
  typedef typename C::Word Word;
  typedef typename C::Store Store;
  constexpr int CHUNKS = C::B / C::VEC;
#pragma unroll
  for (int j = 0; j < (CHUNKS + C::TPU - 1) / C::TPU; ++j) {
    const int ch = tid + j * C::TPU;
    if (CHUNKS % C::TPU != 0 && ch >= CHUNKS) break;
    const int pos = ch * C::VEC;
    const uint4 t = *reinterpret_cast<const uint4 *>(tile + C::taddr(pos));
    if (sizeof(Store) == sizeof(Word)) {
      *reinterpret_cast<uint4 *>(g + pos) = t;
    } else {  // 16-bit limbs: narrow 4 words to 4 limbs (8 bytes)
      const Word *w = reinterpret_cast<const Word *>(&t);
      uint2 o;
      o.x = (uint32_t)w[0] | ((uint32_t)w[1] << 16);
      o.y = (uint32_t)w[2] | ((uint32_t)w[3] << 16);
      *reinterpret_cast<uint2 *>(g + pos) = o;
    }
  }
}
// copy-out with the coefficient-wise product fused in: g[pos] = tile[pos] * other[pos] mod p  (ops.hpp:184-219)
template <class C, int LB> __device__ __forceinline__ void tile_to_gmem_mul(const typename C::Word *tile, typename C::Store *g,
                                                                            const typename C::Store *other, int tid, typename C::Word p,
                                                                            uint64_t k) {
  typedef typename C::Word Word;
  typedef typename C::Store Store;
  constexpr int CHUNKS = C::B / C::VEC;
#pragma unroll
  for (int j = 0; j < (CHUNKS + C::TPU - 1) / C::TPU; ++j) {
    const int ch = tid + j * C::TPU;
    if (CHUNKS % C::TPU != 0 && ch >= CHUNKS) break;
    const int pos = ch * C::VEC;
    uint4 t = *reinterpret_cast<const uint4 *>(tile + C::taddr(pos));
    Word *w = reinterpret_cast<Word *>(&t);
    if (sizeof(Store) == sizeof(Word)) {
      const uint4 o = __ldg(reinterpret_cast<const uint4 *>(other + pos));
      const Word *ow = reinterpret_cast<const Word *>(&o);
#pragma unroll
      for (int i = 0; i < C::VEC; ++i) w[i] = PW<LB>::mulmod(w[i], ow[i], p, k);
      *reinterpret_cast<uint4 *>(g + pos) = t;
    } else {
      const uint2 o = __ldg(reinterpret_cast<const uint2 *>(other + pos));
      const Word ow[4] = {o.x & 0xffffu, o.x >> 16, o.y & 0xffffu, o.y >> 16};
#pragma unroll
      for (int i = 0; i < 4; ++i) w[i] = PW<LB>::mulmod(w[i], ow[i], p, k);
      uint2 r;
      r.x = (uint32_t)w[0] | ((uint32_t)w[1] << 16);
      r.y = (uint32_t)w[2] | ((uint32_t)w[3] << 16);
      *reinterpret_cast<uint2 *>(g + pos) = r;
    }
  }
}
template <class C> __device__ __forceinline__ void gmem_to_tile(typename C::Word *tile, const typename C::Store *g, int tid) {
  typedef typename C::Word Word;
  typedef typename C::Store Store;
  constexpr int CHUNKS = C::B / C::VEC;
#pragma unroll
  for (int j = 0; j < (CHUNKS + C::TPU - 1) / C::TPU; ++j) {
    const int ch = tid + j * C::TPU;
    if (CHUNKS % C::TPU != 0 && ch >= CHUNKS) break;
    const int pos = ch * C::VEC;
    uint4 t;
    if (sizeof(Store) == sizeof(Word)) {
      t = ld_coef(reinterpret_cast<const uint4 *>(g + pos));
    } else {
      const uint2 i = ld_coef(reinterpret_cast<const uint2 *>(g + pos));
      t.x = i.x & 0xffffu; t.y = i.x >> 16; t.z = i.y & 0xffffu; t.w = i.y >> 16;
    }
    *reinterpret_cast<uint4 *>(tile + C::taddr(pos)) = t;
  }
}

// pass-0 register window <-> the unit's slab in global memory (SPLIT == 0: `tid` is the thread's index inside the unit).  For
// fixed k the threads touch consecutive limbs (8 bytes per lane), or with adjacent columns consecutive 16-byte vectors.
template <class C> NFLGPU_DEVFN void window_load(typename C::Word (&x)[C::E], const typename C::Store *unit, int tid) {
  typedef typename C::Word Word;
  if (C::ADJ) {
    constexpr int r0 = plan_r(C::n, C::WB, 0), COLS = 1 << C::CB;
    const typename C::Store *base = unit + (tid << C::CB);
#pragma unroll
    for (int kh = 0; kh < (1 << r0); ++kh) {
#pragma unroll
      for (int v = 0; v < COLS / C::VEC; ++v) {
        const uint4 t = ld_coef(reinterpret_cast<const uint4 *>(base + (kh << (C::n - r0)) + v * C::VEC));
        const Word *w = reinterpret_cast<const Word *>(&t);
#pragma unroll
        for (int j = 0; j < C::VEC; ++j) x[kh * COLS + v * C::VEC + j] = w[j];
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < C::E; ++k) x[k] = (Word)ld_coef(unit + pass_pos<C, C::SPLIT>(tid, k));
  }
}
template <class C> NFLGPU_DEVFN void window_store(const typename C::Word (&x)[C::E], typename C::Store *unit, int tid) {
  typedef typename C::Word Word;
  typedef typename C::Store Store;
  if (C::ADJ) {
    constexpr int r0 = plan_r(C::n, C::WB, 0), COLS = 1 << C::CB;
    Store *base = unit + (tid << C::CB);
#pragma unroll
    for (int kh = 0; kh < (1 << r0); ++kh) {
#pragma unroll
      for (int v = 0; v < COLS / C::VEC; ++v) {
        uint4 t;
        Word *w = reinterpret_cast<Word *>(&t);
#pragma unroll
        for (int j = 0; j < C::VEC; ++j) w[j] = x[kh * COLS + v * C::VEC + j];
        *reinterpret_cast<uint4 *>(base + (kh << (C::n - r0)) + v * C::VEC) = t;
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < C::E; ++k) unit[pass_pos<C, C::SPLIT>(tid, k)] = (Store)x[k];
  }
}

// the same copy issued as cp.async (LDGSTS: global -> shared without passing through registers); completion: cp_async_wait()
template <class C> __device__ __forceinline__ void gmem_to_tile_async(typename C::Word *tile, const typename C::Store *g, int tid) {
  static_assert(sizeof(typename C::Store) == sizeof(typename C::Word), "cp.async copies limbs as they are");
  constexpr int CHUNKS = C::B / C::VEC;
#pragma unroll
  for (int j = 0; j < (CHUNKS + C::TPU - 1) / C::TPU; ++j) {
    const int ch = tid + j * C::TPU;
    if (CHUNKS % C::TPU != 0 && ch >= CHUNKS) break;
    const int pos = ch * C::VEC;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(tile + C::taddr(pos))), "l"(g + pos) : "memory");
  }
}
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- pass chains (compile-time recursion over the passes) ---------------------------------------------------

// forward: passes SPLIT+1 .. NP-1 after pass SPLIT has stored its result into the tile (the caller has synchronised the slot)
template <class C, int PASS> struct FwdChain {
  static __device__ __forceinline__ void run(typename C::Word (&x)[C::E], typename C::Word *tile, const typename C::TW *tw,
                                             typename C::Word p, typename C::Word np, typename C::Word twop, int tid, int slot,
                                             int lane_base) {
    tile_load<C, PASS>(x, tile, tid);
    fwd_pass<C, PASS>(x, pass_tw<C, PASS>(tw, tid), np, twop);
#if !(defined(NFLGPU_ABL) && (NFLGPU_ABL & 8))  // (timing experiment 8: no canonicalisation)
    if (PASS == C::NP - 1) {
#pragma unroll
      for (int k = 0; k < C::E; ++k) x[k] = fwd_canon<C>(x[k], p, twop);
    }
#endif
    tile_store<C, PASS>(x, tile, tid);
    if (PASS + 1 < C::NP) unit_sync<C>(slot, lane_base);
    FwdChain<C, PASS + 1>::run(x, tile, tw, p, np, twop, tid, slot, lane_base);
  }
};
template <class C> struct FwdChain<C, C::NP> {
  static __device__ __forceinline__ void run(typename C::Word (&)[C::E], typename C::Word *, const typename C::TW *, typename C::Word,
                                             typename C::Word, typename C::Word, int, int, int) {}
};

// inverse: passes NP-1 .. SPLIT+1 (tile -> registers -> tile) after the caller has synchronised the slot; pass SPLIT is
// done by the kernel body
template <class C, int PASS> struct InvChain {
  static __device__ __forceinline__ void run(typename C::Word (&x)[C::E], typename C::Word *tile, const typename C::TW *tw,
                                             typename C::Word p, typename C::Word np, typename C::Word twop, const typename C::TW ninv,
                                             int tid, int slot, int lane_base) {
    tile_load<C, PASS>(x, tile, tid);
    inv_pass<C, PASS>(x, pass_tw<C, PASS>(tw, tid), p, np, twop, ninv);
    tile_store<C, PASS>(x, tile, tid);
    if (PASS - 1 > C::SPLIT) unit_sync<C>(slot, lane_base);
    InvChain<C, PASS - 1>::run(x, tile, tw, p, np, twop, ninv, tid, slot, lane_base);
  }
};
template <class C> struct InvChain<C, C::SPLIT> {
  static __device__ __forceinline__ void run(typename C::Word (&)[C::E], typename C::Word *, const typename C::TW *, typename C::Word,
                                             typename C::Word, typename C::Word, const typename C::TW, int, int, int) {}
};

// ---- which sub-block a unit slot works on next ------------------------------------------------------------------

// Static: slot s of CTA `rank` takes sub-blocks rank*SLOTS + s, + ctas_per_residue*SLOTS, ...
// Dynamic: the slot's first thread claims the next index of its residue with one atomicAdd, issued a whole unit ahead
// (the returned value is only consumed at the end of the current unit, so its latency is hidden), and hands it to the
// other threads of the slot through a double-buffered shared-memory word and the slot's barrier.
template <class C> struct UnitWalk {
  uint32_t *cnt;        // this residue's counter
  uint32_t *box;        // this slot's two mailbox words (SLOTS apart)
  uint32_t ahead, cur;  // leader only: index claimed for the next iteration; all: current index
  uint32_t stride;
  int par, slot, lane_base;
  bool leader;
  static __device__ __forceinline__ uint32_t claim(uint32_t *counter) {
    uint32_t v;
    asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(v) : "l"(counter) : "memory");
    return v;
  }
  __device__ __forceinline__ UnitWalk(const NttArgs &a, unsigned char *smem, int cm, int rank, int slot_, int tl, int lane_base_,
                                      bool rearm = false, size_t tw_bytes = C::TW_BYTES)
      : cnt(a.sched + cm), box(reinterpret_cast<uint32_t *>(smem + tw_bytes + 16) + slot_), ahead(0),
        cur((uint32_t)rank * C::SLOTS + slot_), stride(a.ctas_per_residue * C::SLOTS), par(0), slot(slot_), lane_base(lane_base_),
        leader(tl == 0) {
    if (C::DYNAMIC) {
      if (rearm) unit_sync<C>(slot, lane_base);  // (residue hop) every thread of the slot has read the previous walk's last mailbox word
      if (leader) box[0] = claim(cnt);
      unit_sync<C>(slot, lane_base);
      cur = box[0];
      par = C::SLOTS;
    }
  }
  __device__ __forceinline__ uint32_t index() const { return cur; }
  // The index of the NEXT iteration, for prefetching: known at once in the static walk; in the dynamic walk the leader
  // publishes its claim through the mailbox (call publish() before a slot barrier that is late enough for the atomic to have
  // returned, peek_next() after it; advance() rewrites the same value).
  __device__ __forceinline__ void publish() {
    if (C::DYNAMIC && !C::CLAIM_LATE && leader) box[par] = ahead;
  }
  __device__ __forceinline__ uint32_t peek_next() const { return C::DYNAMIC ? (C::CLAIM_LATE ? 0xffffffffu : box[par]) : cur + stride; }
  // call at the top of an iteration
  __device__ __forceinline__ void claim_ahead() {
    if (C::DYNAMIC && !C::CLAIM_LATE && leader) ahead = claim(cnt);
  }
  // call at the bottom of an iteration; in dynamic mode this is also a slot-wide barrier
  __device__ __forceinline__ void advance() {
    if (C::DYNAMIC) {
      if (leader) box[par] = C::CLAIM_LATE ? claim(cnt) : ahead;
      unit_sync<C>(slot, lane_base);
      cur = box[par];
      par ^= C::SLOTS;
    } else {
      cur += stride;
    }
  }
  // after the loop: the last CTA of the grid to get here re-arms the counters for the next launch
  __device__ __forceinline__ static void finish(const NttArgs &a) {
    if (C::DYNAMIC) {
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(a.sched + a.nmoduli, 1u) == gridDim.x - 1) {
          for (uint32_t i = 0; i <= a.nmoduli; ++i) a.sched[i] = 0;
        }
      }
    }
  }
};

// L2 prefetch of the sub-block the slot transforms in its next iteration (measured on B200, profiles/r01e_variants.log:
// -1 .. -4 % per launch; the unit's first loads otherwise wait for HBM with nothing of their own warp to overlap)
template <class C> __device__ __forceinline__ void next_unit_prefetch(const UnitWalk<C> &walk, const typename C::Store *src, uint32_t nmoduli,
                                                                      int cm, uint32_t nblocks, int tl) {
  const uint32_t jn = walk.peek_next();
  if (jn < nblocks) prefetch_unit<C>(src + ((size_t)(jn >> C::LOGG) * nmoduli + cm) * C::N + (size_t)(jn & ((1u << C::LOGG) - 1)) * C::B, tl);
}

// ---- kernels -------------------------------------------------------------------------------------------------

#ifdef NFLGPU_TRACE
// experiment builds only (tools/variants.sh ... -DNFLGPU_TRACE): when does every unit slot of the forward kernel finish?
// [0] = earliest CTA start (globaltimer, ns), [1 + cta * SLOTS + slot] = that slot's finish time
static __device__ unsigned long long nflgpu_trace_buf[1 + 8192];
__device__ __forceinline__ unsigned long long trace_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif

template <class C, bool INV = false> __device__ __forceinline__ const typename C::TW *stage_twiddles(const NttArgs &a, int cm, unsigned char *smem) {
  typedef typename C::TW TW;
  constexpr size_t TWB = INV ? C::TW_BYTES_INV : C::TW_BYTES;
  const TW *twg = reinterpret_cast<const TW *>(a.tw) + (size_t)cm * (INV ? C::INV_TW : C::N);
  if (!C::TW_SMEM) return twg;
  TW *tws = reinterpret_cast<TW *>(smem);
  void *bar = smem + TWB;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, (uint32_t)TWB);
    tma_load_1d(tws, twg, (uint32_t)TWB, bar);
  }
  mbar_wait(bar, 0);
  return tws;
}

template <int LB, int LOGN, bool MUL>
__global__ void __launch_bounds__(NttCfg<LB, LOGN>::THREADS, NttCfg<LB, LOGN>::MIN_BLOCKS) ntt_fwd_kernel(const NttArgs a) {
  typedef NttCfg<LB, LOGN> C;
  typedef typename C::Word Word;
  typedef typename C::Store Store;
  typedef typename C::TW TW;
  constexpr int S = C::SPLIT;
  extern __shared__ __align__(128) unsigned char smem[];
  const int cm0 = blockIdx.x % a.nmoduli, rank = blockIdx.x / a.nmoduli;
  const int slot = threadIdx.x / C::TPU, tl = threadIdx.x % C::TPU;
  const int lane_base = (threadIdx.x & 31) - (tl & 31);
  Word *tile = reinterpret_cast<Word *>(smem + C::TILE_OFF) + (size_t)slot * C::TILE_WORDS;
  const Store *src = reinterpret_cast<const Store *>(a.src);
  Store *dst = reinterpret_cast<Store *>(a.dst);

  // one iteration = one sub-block (= one whole unit when SPLIT == 0)
  const uint32_t nblocks = a.batch << C::LOGG;  // the launcher keeps batch << LOGG below 2^31
#ifdef NFLGPU_TRACE
  if (threadIdx.x == 0) atomicMin(&nflgpu_trace_buf[0], trace_now());
#endif
  for (int hop = 0; hop < (C::HOP ? (int)a.nmoduli : 1); ++hop) {  // (HOP: after its own residue the CTA helps with the others)
  const int cm = cm0 + hop < (int)a.nmoduli ? cm0 + hop : cm0 + hop - (int)a.nmoduli;
  const TW *tw = stage_twiddles<C>(a, cm, smem);
  const Word p = reinterpret_cast<const Word *>(a.moduli)[cm], twop = 2 * p, np = opaque_neg(p);
  UnitWalk<C> walk(a, smem, cm, rank, slot, tl, lane_base, hop > 0);
  if constexpr (C::TMA_SLAB) {  // experiment build only (NttCfg::TMA_SLAB)
    unsigned char *sbar = smem + C::TW_BYTES + 16 + C::SCHED_BYTES + (size_t)slot * 8;
    auto slab_of = [&](uint32_t j) { return src + ((size_t)j * a.nmoduli + cm) * C::N; };
    auto fetch = [&](uint32_t j) {  // elected thread: the slot's tile <- slab j
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the tile's earlier generic-proxy accesses come first
      mbar_expect_tx(sbar, (uint32_t)(C::N * sizeof(Word)));
      tma_load_1d(tile, slab_of(j), (uint32_t)(C::N * sizeof(Word)), sbar);
    };
    if (tl == 0) {
      mbar_init(sbar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unit_sync<C>(slot, lane_base);
    if (tl == 0 && walk.index() < nblocks) fetch(walk.index());
    uint32_t phase = 0;
    for (; walk.index() < nblocks; walk.advance()) {
      const uint32_t j = walk.index();
      const size_t ubase = ((size_t)j * a.nmoduli + cm) * C::N;
      const int tid = tl;
      next_unit_prefetch<C>(walk, src, a.nmoduli, cm, nblocks, tl);  // L2, one unit ahead of the bulk copy
      Word x[C::E];
      mbar_wait(sbar, phase);
      phase ^= 1;
#pragma unroll
      for (int k = 0; k < C::E; ++k) x[k] = tile[pass_pos<C, S>(tid, k)];  // dense slab image: lane-contiguous, conflict free
      unit_sync<C>(slot, lane_base);  // everybody holds its window: the padded layout may overwrite the image
      fwd_pass<C, S>(x, pass_tw<C, S>(tw, tid), np, twop);
      tile_store<C, S>(x, tile, tid);
      unit_sync<C>(slot, lane_base);
      FwdChain<C, S + 1>::run(x, tile, tw, p, np, twop, tid, slot, lane_base);
      unit_sync<C>(slot, lane_base);
      if (MUL) tile_to_gmem_mul<C, LB>(tile, dst + ubase, reinterpret_cast<const Store *>(a.other) + ubase, tl, p, a.consts[cm]);
      else tile_to_gmem<C>(tile, dst + ubase, tl);
      unit_sync<C>(slot, lane_base);  // the copy-out has released the tile
      const uint32_t jn = walk.peek_next();
      if (tl == 0 && jn < nblocks) fetch(jn);
    }
  } else {
  Word x[C::E];
  // pass-0 window of sub-block j, straight from global memory: for fixed k the threads touch consecutive limbs
  auto load_window = [&](uint32_t j) {
    const size_t ub = ((size_t)(j >> C::LOGG) * a.nmoduli + cm) * C::N;
    const int t = (int)((j & ((1u << C::LOGG) - 1)) * C::TPU) + tl;
    window_load<C>(x, src + ub, t);
  };
  if (C::PIPE_FWD && walk.index() < nblocks) load_window(walk.index());
  for (; walk.index() < nblocks; walk.advance()) {
    walk.claim_ahead();
    const uint32_t j = walk.index();
    const uint32_t b = j >> C::LOGG, g = j & ((1u << C::LOGG) - 1);
    const size_t ubase = ((size_t)b * a.nmoduli + cm) * C::N;
    const int tid = (int)(g * C::TPU) + tl;  // index inside the whole unit
    if (!C::DYNAMIC) next_unit_prefetch<C>(walk, src, a.nmoduli, cm, nblocks, tl);  // static walk: the next index is known now
    if (!C::PIPE_FWD) {
#if defined(NFLGPU_ABL) && (NFLGPU_ABL & 1)  // timing experiment: no global loads (results are wrong)
#pragma unroll
      for (int k = 0; k < C::E; ++k) x[k] = (Word)(ubase + pass_pos<C, S>(tid, k)) * 0x9E3779B97F4A7C15ull >> 3;
#else
      window_load<C>(x, src + ubase, tid);
#endif
    }
    fwd_pass<C, S>(x, pass_tw<C, S>(tw, tid), np, twop);
    if (C::NP - S == 1) {
#pragma unroll
      for (int k = 0; k < C::E; ++k) {
        Word v = fwd_canon<C>(x[k], p, twop);
        if (MUL) v = PW<LB>::mulmod(v, (Word)__ldg(reinterpret_cast<const Store *>(a.other) + ubase + pass_pos<C, S>(tid, k)), p, a.consts[cm]);
        dst[ubase + pass_pos<C, S>(tid, k)] = (Store)v;
      }
    } else {
      const size_t bbase = ubase + (size_t)g * C::B;
      if (!C::DYNAMIC) unit_sync<C>(slot, lane_base);  // previous sub-block's copy-out has finished reading the tile (advance() syncs in dynamic mode)
      tile_store<C, S>(x, tile, tid);
      walk.publish();
      unit_sync<C>(slot, lane_base);
      const uint32_t jn = C::PIPE_FWD ? walk.peek_next() : 0xffffffffu;
      if (C::DYNAMIC) next_unit_prefetch<C>(walk, src, a.nmoduli, cm, nblocks, tl);  // the leader's claim has just been published
      FwdChain<C, S + 1>::run(x, tile, tw, p, np, twop, tid, slot, lane_base);
      unit_sync<C>(slot, lane_base);
      if (C::PIPE_FWD && jn < nblocks) load_window(jn);  // (PIPE_FWD) in flight while the tile is copied out
#if defined(NFLGPU_ABL) && (NFLGPU_ABL & 2)  // timing experiment: one store per unit instead of the copy-out
      if (tl == 0) dst[bbase] = (Store)tile[0];
#else
      if (MUL) tile_to_gmem_mul<C, LB>(tile, dst + bbase, reinterpret_cast<const Store *>(a.other) + bbase, tl, p, a.consts[cm]);
      else tile_to_gmem<C>(tile, dst + bbase, tl);
#endif
    }
  }
  }  // (regular path)
  }  // hop
#ifdef NFLGPU_TRACE
  if (tl == 0 && blockIdx.x * C::SLOTS + slot < 8192) nflgpu_trace_buf[1 + blockIdx.x * C::SLOTS + slot] = trace_now();
#endif
  UnitWalk<C>::finish(a);
}

template <int LB, int LOGN>
__global__ void __launch_bounds__(NttCfg<LB, LOGN>::THREADS, NttCfg<LB, LOGN>::MIN_BLOCKS) ntt_inv_kernel(const NttArgs a) {
  typedef NttCfg<LB, LOGN> C;
  typedef typename C::Word Word;
  typedef typename C::Store Store;
  typedef typename C::TW TW;
  constexpr int S = C::SPLIT;
  extern __shared__ __align__(128) unsigned char smem[];
  const int cm0 = blockIdx.x % a.nmoduli, rank = blockIdx.x / a.nmoduli;
  const int slot = threadIdx.x / C::TPU, tl = threadIdx.x % C::TPU;
  const int lane_base = (threadIdx.x & 31) - (tl & 31);
  Word *tile = reinterpret_cast<Word *>(smem + C::TILE_OFF_INV) + (size_t)slot * C::TILE_WORDS;
  const Store *src = reinterpret_cast<const Store *>(a.src);
  Store *dst = reinterpret_cast<Store *>(a.dst);

  const uint32_t nblocks = a.batch << C::LOGG;  // the launcher keeps batch << LOGG below 2^31
  const int cm = cm0;  // (the inverse kernels stay bound to their residue: see NttCfg::HOP)
  const TW *tw = stage_twiddles<C, true>(a, cm, smem);
  const TW ninv = tw[C::N - 1];
  const Word p = reinterpret_cast<const Word *>(a.moduli)[cm], twop = 2 * p, np = opaque_neg(p);
  UnitWalk<C> walk(a, smem, cm, rank, slot, tl, lane_base, false, C::TW_BYTES_INV);
  auto slab_of = [&](uint32_t j) { return src + ((size_t)(j >> C::LOGG) * a.nmoduli + cm) * C::N + (size_t)(j & ((1u << C::LOGG) - 1)) * C::B; };
  if constexpr (C::PIPE_INV) {
    if (walk.index() < nblocks) gmem_to_tile_async<C>(tile, slab_of(walk.index()), tl);
  }
  for (; walk.index() < nblocks; walk.advance()) {
    walk.claim_ahead();
    const uint32_t j = walk.index();
    const uint32_t b = j >> C::LOGG, g = j & ((1u << C::LOGG) - 1);
    const size_t ubase = ((size_t)b * a.nmoduli + cm) * C::N;
    const int tid = (int)(g * C::TPU) + tl;
    if (!C::DYNAMIC && !C::PIPE_INV) next_unit_prefetch<C>(walk, src, a.nmoduli, cm, nblocks, tl);
    Word x[C::E];
    if (C::NP - S == 1) {
#pragma unroll
      for (int k = 0; k < C::E; ++k) x[k] = (Word)ld_coef(src + ubase + pass_pos<C, S>(tid, k));
    } else if constexpr (C::PIPE_INV) {
      cp_async_wait();  // this thread's part of the sub-block (issued during the previous iteration) has landed
      walk.publish();
      unit_sync<C>(slot, lane_base);  // ... and everybody else's; the leader's claim is visible
      const uint32_t jn = walk.peek_next();
      InvChain<C, C::NP - 1>::run(x, tile, tw, p, np, twop, ninv, tid, slot, lane_base);
      unit_sync<C>(slot, lane_base);
      tile_load<C, S>(x, tile, tid);
      unit_sync<C>(slot, lane_base);  // every thread holds its last-pass window: the tile is free for the next sub-block
      if (jn < nblocks) gmem_to_tile_async<C>(tile, slab_of(jn), tl);  // lands while the last pass computes and stores
    } else {
      if (!C::DYNAMIC) unit_sync<C>(slot, lane_base);  // previous sub-block's last pass has finished reading the tile (advance() syncs in dynamic mode)
      gmem_to_tile<C>(tile, src + ubase + (size_t)g * C::B, tl);
      walk.publish();
      unit_sync<C>(slot, lane_base);
      if (C::DYNAMIC) next_unit_prefetch<C>(walk, src, a.nmoduli, cm, nblocks, tl);
      InvChain<C, C::NP - 1>::run(x, tile, tw, p, np, twop, ninv, tid, slot, lane_base);
      unit_sync<C>(slot, lane_base);
      tile_load<C, S>(x, tile, tid);
    }
    inv_pass<C, S>(x, pass_tw<C, S>(tw, tid), p, np, twop, ninv);
    // this pass writes straight to global memory (lane-contiguous for fixed k)
    window_store<C>(x, dst + ubase, tid);
  }
  UnitWalk<C>::finish(a);
}

// Global-memory pass PASS (< SPLIT) of a split transform: every thread owns the E coefficients of one butterfly
// network of this pass (they are 2^c words apart), so the pass needs no exchange: registers <-> HBM, in place on dst.
template <int LB, int LOGN, int PASS, bool INV>
__global__ void __launch_bounds__(256) ntt_gpass_kernel(const NttArgs a) {
  typedef NttCfg<LB, LOGN> C;
  typedef typename C::Word Word;
  typedef typename C::Store Store;
  typedef typename C::TW TW;
  constexpr int LOGT = C::n - C::e;  // threads per unit
  const Store *src = reinterpret_cast<const Store *>(a.src);
  Store *dst = reinterpret_cast<Store *>(a.dst);
  const uint64_t total = ((uint64_t)a.batch * a.nmoduli) << LOGT;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t u = t >> LOGT;
    const int tid = (int)(t & ((1u << LOGT) - 1));
    const int cm = (int)(u % a.nmoduli);
    const TW *tw = reinterpret_cast<const TW *>(a.tw) + (size_t)cm * (INV ? C::INV_TW : C::N);
    const Word p = reinterpret_cast<const Word *>(a.moduli)[cm], twop = 2 * p, np = opaque_neg(p);
    const size_t ubase = (size_t)u * C::N;
    Word x[C::E];
#pragma unroll
    for (int k = 0; k < C::E; ++k) x[k] = (Word)ld_coef(src + ubase + pass_pos<C, PASS>(tid, k));
    if (INV) inv_pass<C, PASS>(x, pass_tw<C, PASS>(tw, tid), p, np, twop, __ldg(tw + C::N - 1));
    else fwd_pass<C, PASS>(x, pass_tw<C, PASS>(tw, tid), np, twop);
#pragma unroll
    for (int k = 0; k < C::E; ++k) dst[ubase + pass_pos<C, PASS>(tid, k)] = (Store)x[k];
  }
}

}  // namespace nflgpu
#endif

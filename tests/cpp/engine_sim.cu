// TEST INFRASTRUCTURE (CPU, no GPU needed): runs the butterfly networks of the NTT kernels on the host.
//
// nfllib_b200/csrc/ntt_engine.cuh marks fwd_pass / inv_pass / pass_pos / pass_tw / fwd_canon and the arithmetic of
// modarith.cuh __host__ __device__; this file executes exactly those functions, pass by pass and "thread" by "thread",
// over a plain array that stands in for the shared-memory tile / HBM slab, with the twiddle tables the product builds
// (tables.cpp).  What it checks, on the CPU suite, is everything of the kernels that is arithmetic or index algebra:
// the pass plan, the lazy ranges (including the 64-bit "top-bit" forward scheme), the table layout and
// the canonicalisation.  What it cannot check — launch geometry, tile padding, barriers, TMA — is covered by the GPU suite.
// The coefficient-wise functors of the pointwise kernels (modmul.cuh Functor<LB, OP>::apply, also used by the fused
// forward-NTT * operand epilogue) are exposed the same way.  Results are compared with the oracle by tests/test_engine_sim.py.
#include "../../nfllib_b200/csrc/host_common.hpp"
#include "../../nfllib_b200/csrc/ntt_engine.cuh"

#include <cstring>
#include <vector>

using namespace nflgpu;

namespace {

template <class C, int PASS> struct SimFwd {
  static void run(typename C::Word *d, const typename C::TW *tw, typename C::Word p) {
    typedef typename C::Word Word;
    const Word np = opaque_neg(p), twop = 2 * p;
    for (int tid = 0; tid < (C::N >> C::e); ++tid) {
      Word x[C::E];
      for (int k = 0; k < C::E; ++k) x[k] = d[pass_pos<C, PASS>(tid, k)];
      fwd_pass<C, PASS>(x, pass_tw<C, PASS>(tw, tid), np, twop);
      if (PASS == C::NP - 1)
        for (int k = 0; k < C::E; ++k) x[k] = fwd_canon<C>(x[k], p, twop);
      for (int k = 0; k < C::E; ++k) d[pass_pos<C, PASS>(tid, k)] = x[k];
    }
    SimFwd<C, PASS + 1>::run(d, tw, p);
  }
};
template <class C> struct SimFwd<C, C::NP> {
  static void run(typename C::Word *, const typename C::TW *, typename C::Word) {}
};

template <class C, int PASS> struct SimInv {
  static void run(typename C::Word *d, const typename C::TW *tw, typename C::Word p) {
    typedef typename C::Word Word;
    const Word np = opaque_neg(p), twop = 2 * p;
    const typename C::TW ninv = tw[C::N - 1];
    for (int tid = 0; tid < (C::N >> C::e); ++tid) {
      Word x[C::E];
      for (int k = 0; k < C::E; ++k) x[k] = d[pass_pos<C, PASS>(tid, k)];
      inv_pass<C, PASS>(x, pass_tw<C, PASS>(tw, tid), p, np, twop, ninv);
      for (int k = 0; k < C::E; ++k) d[pass_pos<C, PASS>(tid, k)] = x[k];
    }
    SimInv<C, PASS - 1>::run(d, tw, p);
  }
};
template <class C> struct SimInv<C, -1> {
  static void run(typename C::Word *, const typename C::TW *, typename C::Word) {}
};

// The same transforms with the exchange between passes going through the kernels' own tile layout (tile_store / tile_load /
// C::taddr: padded rows or the XOR swizzle), "thread" by "thread", the way ntt_fwd_kernel / ntt_inv_kernel use the tile:
// forward = global -> registers -> pass S -> tile ... last pass -> tile -> 16-byte copy-out; inverse = 16-byte copy-in -> tile ->
// passes NP-1 .. S+1 in the tile -> pass S from the tile -> global.  Single-tile shapes with more than one pass only.
template <class C, int PASS> struct SimFwdTile {
  static void run(typename C::Word *tile, const typename C::TW *tw, typename C::Word p) {
    typedef typename C::Word Word;
    const Word np = opaque_neg(p), twop = 2 * p;
    for (int tid = 0; tid < C::TPU; ++tid) {
      Word x[C::E];
      tile_load<C, PASS>(x, tile, tid);
      fwd_pass<C, PASS>(x, pass_tw<C, PASS>(tw, tid), np, twop);
      if (PASS == C::NP - 1)
        for (int k = 0; k < C::E; ++k) x[k] = fwd_canon<C>(x[k], p, twop);
      tile_store<C, PASS>(x, tile, tid);
    }
    SimFwdTile<C, PASS + 1>::run(tile, tw, p);
  }
};
template <class C> struct SimFwdTile<C, C::NP> {
  static void run(typename C::Word *, const typename C::TW *, typename C::Word) {}
};
template <class C, int PASS> struct SimInvTile {
  static void run(typename C::Word *tile, const typename C::TW *tw, typename C::Word p) {
    typedef typename C::Word Word;
    const Word np = opaque_neg(p), twop = 2 * p;
    const typename C::TW ninv = tw[C::N - 1];
    for (int tid = 0; tid < C::TPU; ++tid) {
      Word x[C::E];
      tile_load<C, PASS>(x, tile, tid);
      inv_pass<C, PASS>(x, pass_tw<C, PASS>(tw, tid), p, np, twop, ninv);
      tile_store<C, PASS>(x, tile, tid);
    }
    SimInvTile<C, PASS - 1>::run(tile, tw, p);
  }
};
template <class C> struct SimInvTile<C, 0> {
  static void run(typename C::Word *, const typename C::TW *, typename C::Word) {}
};

template <int LB, int LOGN> int sim_tile(int inverse, uint64_t p, uint64_t root, uint64_t kmax, uint64_t *data) {
  typedef NttCfg<LB, LOGN> C;
  typedef typename C::Word Word;
  typedef typename C::TW TW;
  if (C::SPLIT != 0 || C::NP < 2) return -1;
  ResidueTables t;
  build_residue_tables(LB, C::WB, C::N, p, root, kmax, &t, false);
  const int entries = inverse ? C::INV_TW : C::N;  // (the inverse table has a second, N^-1-scaled half when C::FOLD)
  std::vector<TW> tw(entries);
  for (int i = 0; i < entries; ++i) {
    tw[i].x = (Word)(inverse ? t.inv_w[i] : t.fwd_w[i]);
    tw[i].y = (Word)(inverse ? t.inv_ws[i] : t.fwd_ws[i]);
  }
  std::vector<Word> tile_store_buf(C::TILE_WORDS + 4, (Word)0xdeadbeef);
  Word *tile = reinterpret_cast<Word *>(((uintptr_t)tile_store_buf.data() + 15) & ~(uintptr_t)15);  // 16-byte vectors
  const Word np = opaque_neg((Word)p), twop = 2 * (Word)p;
  std::vector<typename C::Store> slab(C::N + 2);  // the unit as it lies in global memory (limbs, 16-byte aligned like a batch buffer)
  for (int i = 0; i < C::N; ++i) slab[i] = (typename C::Store)data[i];
  if (!inverse) {
    for (int tid = 0; tid < C::TPU; ++tid) {
      Word x[C::E];
      window_load<C>(x, slab.data(), tid);  // the kernels' own pass-0 global-memory access (8-byte columns or 16-byte vectors)
      fwd_pass<C, 0>(x, pass_tw<C, 0>(tw.data(), tid), np, twop);
      tile_store<C, 0>(x, tile, tid);
    }
    SimFwdTile<C, 1>::run(tile, tw.data(), (Word)p);
    for (int ch = 0; ch < C::B / C::VEC; ++ch)  // tile_to_gmem
      for (int j = 0; j < C::VEC; ++j) data[ch * C::VEC + j] = (uint64_t)tile[C::taddr(ch * C::VEC) + j];
  } else {
    for (int ch = 0; ch < C::B / C::VEC; ++ch)  // gmem_to_tile
      for (int j = 0; j < C::VEC; ++j) tile[C::taddr(ch * C::VEC) + j] = (Word)data[ch * C::VEC + j];
    SimInvTile<C, C::NP - 1>::run(tile, tw.data(), (Word)p);
    const TW ninv = tw[C::N - 1];
    for (int tid = 0; tid < C::TPU; ++tid) {
      Word x[C::E];
      tile_load<C, 0>(x, tile, tid);
      inv_pass<C, 0>(x, pass_tw<C, 0>(tw.data(), tid), (Word)p, np, twop, ninv);
      window_store<C>(x, slab.data(), tid);
    }
    for (int i = 0; i < C::N; ++i) data[i] = (uint64_t)slab[i];
  }
  return 0;
}

template <int LB, int LOGN> int sim_one(int inverse, uint64_t p, uint64_t root, uint64_t kmax, uint64_t *data) {
  typedef NttCfg<LB, LOGN> C;
  typedef typename C::Word Word;
  typedef typename C::TW TW;
  ResidueTables t;
  build_residue_tables(LB, C::WB, C::N, p, root, kmax, &t, false);
  const int entries = inverse ? C::INV_TW : C::N;  // (the inverse table has a second, N^-1-scaled half when C::FOLD)
  std::vector<TW> tw(entries);
  for (int i = 0; i < entries; ++i) {
    tw[i].x = (Word)(inverse ? t.inv_w[i] : t.fwd_w[i]);
    tw[i].y = (Word)(inverse ? t.inv_ws[i] : t.fwd_ws[i]);
  }
  std::vector<Word> d(C::N);
  for (int i = 0; i < C::N; ++i) d[i] = (Word)data[i];
  if (inverse) SimInv<C, C::NP - 1>::run(d.data(), tw.data(), (Word)p);
  else SimFwd<C, 0>::run(d.data(), tw.data(), (Word)p);
  for (int i = 0; i < C::N; ++i) data[i] = (uint64_t)d[i];
  return 0;
}

}  // namespace

template <int LB, int OP> void pw_all(uint64_t p, uint64_t k, const uint64_t *a, const uint64_t *b, const uint64_t *c, const uint64_t *d,
                                      uint64_t *out, size_t n) {
  typedef typename PW<LB>::Word Word;
  for (size_t i = 0; i < n; ++i)
    out[i] = (uint64_t)Functor<LB, OP>::apply((Word)a[i], b ? (Word)b[i] : 0, c ? (Word)c[i] : 0, d ? (Word)d[i] : 0, (Word)p, k);
}
template <int LB> int pw_dispatch(int op, uint64_t p, uint64_t k, const uint64_t *a, const uint64_t *b, const uint64_t *c, const uint64_t *d,
                                  uint64_t *out, size_t n) {
  switch (op) {
    case PW_ADD: pw_all<LB, PW_ADD>(p, k, a, b, c, d, out, n); return 0;
    case PW_SUB: pw_all<LB, PW_SUB>(p, k, a, b, c, d, out, n); return 0;
    case PW_MUL: pw_all<LB, PW_MUL>(p, k, a, b, c, d, out, n); return 0;
    case PW_MUL_SHOUP: pw_all<LB, PW_MUL_SHOUP>(p, k, a, b, c, d, out, n); return 0;
    case PW_COMPUTE_SHOUP: pw_all<LB, PW_COMPUTE_SHOUP>(p, k, a, b, c, d, out, n); return 0;
    case PW_MULADD: pw_all<LB, PW_MULADD>(p, k, a, b, c, d, out, n); return 0;
    case PW_MULADD_SHOUP: pw_all<LB, PW_MULADD_SHOUP>(p, k, a, b, c, d, out, n); return 0;
  }
  return -1;
}

#define SIM_CASE(LB, LOGN) \
  case LOGN: return sim_one<LB, LOGN>(inverse, p, root, kmax, data);

// data: N residues of one (polynomial, modulus) unit, widened to uint64_t, transformed in place.
// Returns 0, or -1 for a size this simulation is not instantiated for.
extern "C" int nflsim_ntt(int limb_bits, int log2_degree, int inverse, uint64_t p, uint64_t root, uint64_t kmax, uint64_t *data) {
  if (limb_bits == 64) {
    switch (log2_degree) {
      SIM_CASE(64, 2) SIM_CASE(64, 3) SIM_CASE(64, 4) SIM_CASE(64, 5) SIM_CASE(64, 6) SIM_CASE(64, 7) SIM_CASE(64, 8)
      SIM_CASE(64, 9) SIM_CASE(64, 10) SIM_CASE(64, 11) SIM_CASE(64, 12) SIM_CASE(64, 13) SIM_CASE(64, 14) SIM_CASE(64, 15)
      SIM_CASE(64, 16) SIM_CASE(64, 17)
    }
  } else if (limb_bits == 32) {
    switch (log2_degree) {
      SIM_CASE(32, 3) SIM_CASE(32, 4) SIM_CASE(32, 5) SIM_CASE(32, 6) SIM_CASE(32, 7) SIM_CASE(32, 8) SIM_CASE(32, 9)
      SIM_CASE(32, 10) SIM_CASE(32, 11) SIM_CASE(32, 12) SIM_CASE(32, 13) SIM_CASE(32, 14) SIM_CASE(32, 15)
    }
  } else if (limb_bits == 16) {
    switch (log2_degree) {
      SIM_CASE(16, 4) SIM_CASE(16, 5) SIM_CASE(16, 6) SIM_CASE(16, 7) SIM_CASE(16, 8) SIM_CASE(16, 9)
    }
  }
  return -1;
}

#undef SIM_CASE
#define SIM_CASE(LB, LOGN) \
  case LOGN: return sim_tile<LB, LOGN>(inverse, p, root, kmax, data);
// As nflsim_ntt, with the passes exchanging through the kernels' shared-memory tile layout (returns -1 for one-pass and split shapes).
extern "C" int nflsim_ntt_tile(int limb_bits, int log2_degree, int inverse, uint64_t p, uint64_t root, uint64_t kmax, uint64_t *data) {
  if (limb_bits == 64) {
    switch (log2_degree) {
      SIM_CASE(64, 6) SIM_CASE(64, 7) SIM_CASE(64, 8) SIM_CASE(64, 9) SIM_CASE(64, 10) SIM_CASE(64, 11) SIM_CASE(64, 12)
      SIM_CASE(64, 13) SIM_CASE(64, 14)
    }
  } else if (limb_bits == 32) {
    switch (log2_degree) {
      SIM_CASE(32, 7) SIM_CASE(32, 8) SIM_CASE(32, 9) SIM_CASE(32, 10) SIM_CASE(32, 11) SIM_CASE(32, 12) SIM_CASE(32, 13)
      SIM_CASE(32, 14) SIM_CASE(32, 15)
    }
  } else if (limb_bits == 16) {
    switch (log2_degree) {
      SIM_CASE(16, 7) SIM_CASE(16, 8) SIM_CASE(16, 9)
    }
  }
  return -1;
}

// One pointwise functor over n coefficients of one residue (operands widened to uint64_t; unused operands may be null).
// op: PwOp of pointwise.h; the per-modulus constant is derived exactly as nflgpu_ctx_create derives it.
extern "C" int nflsim_pointwise(int limb_bits, int op, uint64_t p, const uint64_t *a, const uint64_t *b, const uint64_t *c, const uint64_t *d,
                                uint64_t *out, size_t n) {
  if (limb_bits == 64) return pw_dispatch<64>(op, p, newton_pn(64, p), a, b, c, d, out, n);
  if (limb_bits == 32) return pw_dispatch<32>(op, p, (uint64_t)((((unsigned __int128)1) << 64) / p), a, b, c, d, out, n);
  if (limb_bits == 16) return pw_dispatch<16>(op, p, 0, a, b, c, d, out, n);
  return -1;
}

// Compile-time expression programs: one kernel per postfix program of eval_shapes.inc.
//
// The reference's expression templates (ops::expr ops.hpp:52-97, _make_op ops.hpp:249-277, evaluator core.hpp:24-37)
// compile every right-hand side such as `a + b*c` into its own fused loop.  The device equivalent for the small trees user
// code actually writes: the program is a template parameter pack, so the token loop unrolls, the value stack lives in
// registers, all leaves are loaded up front (memory-level parallelism) and the kernel is as HBM-bound as a hand-written
// functor kernel.  nflgpu_eval canonicalises a caller's program (leaves renumbered by first appearance) and looks it up here;
// anything not in the table runs on the interpreter (eval_kernel, pointwise.cu).
#include "pointwise.h"
#include "modarith.cuh"
#include "modmul.cuh"
#include "vecio.cuh"

namespace nflgpu {

template <uint8_t... PROG> struct EvProgram {
  static constexpr int NT = sizeof...(PROG);
  static __host__ __device__ constexpr int tok(int i) {
    constexpr uint8_t p[NT] = {PROG...};
    return p[i];
  }
  static __host__ __device__ constexpr int nleaves() {
    int n = 0;
    for (int i = 0; i < NT; ++i) if (tok(i) < EV_MAX_OPERANDS && tok(i) + 1 > n) n = tok(i) + 1;
    return n;
  }
  // stack pointer before token i
  static __host__ __device__ constexpr int sp(int i) {
    int s = 0;
    for (int t = 0; t < i; ++t) s += tok(t) < EV_MAX_OPERANDS ? 1 : (tok(t) == EV_MUL_SHOUP ? -2 : (tok(t) == EV_COMPUTE_SHOUP ? 0 : -1));
    return s;
  }
  static __host__ __device__ constexpr int depth() {
    int d = 0;
    for (int i = 0; i <= NT; ++i) if (sp(i) > d) d = sp(i);
    return d;
  }
};

template <int LB, uint8_t... PROG>
__global__ void __launch_bounds__(256) eval_static_kernel(const EvArgs a) {
  typedef EvProgram<PROG...> P;
  typedef typename PW<LB>::Word Word;
  typedef typename PW<LB>::Store Store;
  constexpr int VEC = PW<LB>::VEC, NL = P::nleaves(), DEPTH = P::depth();
  const uint32_t cm = blockIdx.y;
  const Word p = (Word)a.moduli[cm];
  const uint64_t kc = a.consts[cm];
  const uint32_t vec_per_row = a.degree / VEC, row_shift = a.log2_degree - (VEC == 2 ? 1 : VEC == 4 ? 2 : 3);
  const uint64_t total = (uint64_t)a.batch * vec_per_row;
  Store *dst = reinterpret_cast<Store *>(a.dst);
  for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t b = v >> row_shift, off = v & (vec_per_row - 1);
    const size_t at = ((size_t)b * a.nmoduli + cm) * a.degree + off * VEC;
    Word leaf[NL][VEC];
#pragma unroll
    for (int l = 0; l < NL; ++l) VecIO<LB>::load(leaf[l], reinterpret_cast<const Store *>(a.operands[l]) + at);
    Word st[DEPTH][VEC];
#pragma unroll
    for (int t = 0; t < P::NT; ++t) {
      const int tok = P::tok(t), sp = P::sp(t);  // compile-time constants once the loop is unrolled
      if (tok < EV_MAX_OPERANDS) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) st[sp][i] = leaf[tok][i];
      } else if (tok == EV_COMPUTE_SHOUP) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) st[sp - 1][i] = Functor<LB, PW_COMPUTE_SHOUP>::apply(st[sp - 1][i], 0, 0, 0, p, kc);
      } else if (tok == EV_MUL_SHOUP) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) st[sp - 3][i] = Functor<LB, PW_MUL_SHOUP>::apply(st[sp - 3][i], st[sp - 2][i], st[sp - 1][i], 0, p, kc);
      } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const Word x = st[sp - 2][i], y = st[sp - 1][i];
          st[sp - 2][i] = tok == EV_ADD ? Functor<LB, PW_ADD>::apply(x, y, 0, 0, p, kc)
                        : tok == EV_SUB ? Functor<LB, PW_SUB>::apply(x, y, 0, 0, p, kc)
                                        : Functor<LB, PW_MUL>::apply(x, y, 0, 0, p, kc);
        }
      }
    }
    VecIO<LB>::store(dst + at, st[0]);
  }
}

template <int LB, uint8_t... PROG> static cudaError_t launch_static(const EvArgs &a, int num_sms, cudaStream_t stream) {
  constexpr int VEC = PW<LB>::VEC;
  const uint64_t total = (uint64_t)a.batch * (a.degree / VEC);
  if (total == 0) return cudaSuccess;
  uint64_t blocks = (total + 255) / 256;
  const uint64_t cap = (uint64_t)num_sms * 8 / a.nmoduli + 1;
  if (blocks > cap) blocks = cap;
  eval_static_kernel<LB, PROG...><<<dim3((unsigned)blocks, a.nmoduli), 256, 0, stream>>>(a);
  return cudaGetLastError();
}

template <int LB> static bool launch_static_limb(uint64_t key, const EvArgs &a, int num_sms, cudaStream_t stream, cudaError_t *err) {
  switch (key) {
#define NFLGPU_EVAL_SHAPE(KEY, ...) case KEY: *err = launch_static<LB, __VA_ARGS__>(a, num_sms, stream); return true;
#include "eval_shapes.inc"
#undef NFLGPU_EVAL_SHAPE
  }
  return false;
}

// true when `key` (sum over tokens of (token + 1) << 8*i, leaves numbered by first appearance) has a compiled kernel; *err
// then holds the launch status
bool launch_eval_static(int limb_bits, uint64_t key, const EvArgs &a, int num_sms, cudaStream_t stream, cudaError_t *err) {
  switch (limb_bits) {
    case 64: return launch_static_limb<64>(key, a, num_sms, stream, err);
    case 32: return launch_static_limb<32>(key, a, num_sms, stream, err);
    case 16: return launch_static_limb<16>(key, a, num_sms, stream, err);
  }
  return false;
}

}  // namespace nflgpu

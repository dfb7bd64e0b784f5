// Coefficient-wise modular functors over device-resident batches (HBM-bound elementwise kernels, sm_100a).
//
// Replaces the reference's expression evaluator loop (core.hpp:24-37) applied to
//   addmod ops.hpp:124-135 | submod ops.hpp:141-151 | mulmod ops.hpp:184-219 | mulmod_shoup ops.hpp:225-242
//   compute_shoup ops.hpp:165-177 | muladd opt/ops.hpp:9-48 | muladd_shoup opt/ops.hpp:56-78
// and their SSE/AVX2 specialisations (sse.hpp:75-153,309-490; avx2.hpp:69-147,311-423).
// One 16-byte vector per thread per operand, grid.y = residue so the modulus is uniform per block and no
// integer division is needed to find it.  Results are canonical and bit-identical to the reference's.
#include "pointwise.h"
#include "modarith.cuh"
#include "modmul.cuh"
#include "vecio.cuh"

namespace nflgpu {

// (the functors themselves are in modmul.cuh: Functor<LB, OP>::apply, shared with the host simulation of the CPU suite)

template <int LB, int OP, int NIN>
__global__ void __launch_bounds__(256) pointwise_kernel(const PwArgs a) {
  typedef typename PW<LB>::Word Word;
  typedef typename PW<LB>::Store Store;
  constexpr int VEC = PW<LB>::VEC;
  const uint32_t cm = blockIdx.y;
  const Word p = (Word)a.moduli[cm];
  const uint64_t k = a.consts[cm];
  const uint32_t vec_per_row = a.degree / VEC, row_shift = a.log2_degree - (VEC == 2 ? 1 : VEC == 4 ? 2 : 3);
  const uint64_t total = (uint64_t)a.batch * vec_per_row;
  Store *dst = reinterpret_cast<Store *>(a.dst);
  const Store *pa = reinterpret_cast<const Store *>(a.a), *pb = reinterpret_cast<const Store *>(a.b);
  const Store *pc = reinterpret_cast<const Store *>(a.c), *pd = reinterpret_cast<const Store *>(a.d);
  for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t b = v >> row_shift, off = v & (vec_per_row - 1);
    const size_t at = ((size_t)b * a.nmoduli + cm) * a.degree + off * VEC;
    Word wa[VEC], wb[VEC], wc[VEC], wd[VEC], wo[VEC];
    VecIO<LB>::load(wa, pa + at);
    if (NIN >= 2) VecIO<LB>::load(wb, pb + at);
    if (NIN >= 3) VecIO<LB>::load(wc, pc + at);
    if (NIN >= 4) VecIO<LB>::load(wd, pd + at);
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      wo[i] = Functor<LB, OP>::apply(wa[i], NIN >= 2 ? wb[i] : 0, NIN >= 3 ? wc[i] : 0, NIN >= 4 ? wd[i] : 0, p, k);
    VecIO<LB>::store(dst + at, wo);
  }
}

template <int LB, int OP, int NIN> static cudaError_t launch_one(const PwArgs &a, int num_sms, cudaStream_t stream) {
  constexpr int VEC = PW<LB>::VEC;
  const uint64_t total = (uint64_t)a.batch * (a.degree / VEC);
  if (total == 0) return cudaSuccess;
  uint64_t blocks = (total + 255) / 256;
  const uint64_t cap = (uint64_t)num_sms * 8 / a.nmoduli + 1;  // ~8 resident blocks of 256 threads per SM over all residues
  if (blocks > cap) blocks = cap;
  dim3 grid((unsigned)blocks, a.nmoduli);
  pointwise_kernel<LB, OP, NIN><<<grid, 256, 0, stream>>>(a);
  return cudaGetLastError();
}

template <int LB> static cudaError_t launch_limb(int op, const PwArgs &a, int num_sms, cudaStream_t s) {
  switch (op) {
    case PW_ADD: return launch_one<LB, PW_ADD, 2>(a, num_sms, s);
    case PW_SUB: return launch_one<LB, PW_SUB, 2>(a, num_sms, s);
    case PW_MUL: return launch_one<LB, PW_MUL, 2>(a, num_sms, s);
    case PW_MUL_SHOUP: return launch_one<LB, PW_MUL_SHOUP, 3>(a, num_sms, s);
    case PW_COMPUTE_SHOUP: return launch_one<LB, PW_COMPUTE_SHOUP, 1>(a, num_sms, s);
    case PW_MULADD: return launch_one<LB, PW_MULADD, 3>(a, num_sms, s);
    case PW_MULADD_SHOUP: return launch_one<LB, PW_MULADD_SHOUP, 4>(a, num_sms, s);
  }
  return cudaErrorInvalidValue;
}

// ---- fused expression evaluator ---------------------------------------------------------------------------------
// The reference evaluates a whole expression tree per coefficient in one loop (ops::expr::load recursion, ops.hpp:69-79,
// driven by core.hpp:24-37).  Here the tree arrives as a postfix program that every thread interprets on a small
// value stack; the program is uniform across the grid, so the interpreter's branches never diverge and the kernel
// stays HBM-bound: each operand is read exactly once and the result written once, whatever the tree.
// The operand file and value stack are indexed by program-dependent values, so the compiler keeps the stack in
// (L1-resident) local memory at 32 registers per thread = full occupancy, which hides the operand loads.  Measured on
// B200 for `a + b*c` over 128 MiB operands (tools/kbench_all.py): 166 us, against 91 us for the specialised muladd kernel
// and 149 us for two separate kernels; variants with a register-resident stack (static-depth dispatch) or with all
// operands prefetched were slower (205 / 216 us: more registers, fewer resident warps).  The interpreter's value is one
// launch and no temporaries for arbitrary trees; the shapes the reference names keep their specialised kernels.
template <int LB>
__global__ void __launch_bounds__(256) eval_kernel(const EvArgs a) {
  typedef typename PW<LB>::Word Word;
  typedef typename PW<LB>::Store Store;
  constexpr int VEC = PW<LB>::VEC;
  const uint32_t cm = blockIdx.y;
  const Word p = (Word)a.moduli[cm];
  const uint64_t kc = a.consts[cm];
  const uint32_t vec_per_row = a.degree / VEC, row_shift = a.log2_degree - (VEC == 2 ? 1 : VEC == 4 ? 2 : 3);
  const uint64_t total = (uint64_t)a.batch * vec_per_row;
  Store *dst = reinterpret_cast<Store *>(a.dst);
  for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t b = v >> row_shift, off = v & (vec_per_row - 1);
    const size_t at = ((size_t)b * a.nmoduli + cm) * a.degree + off * VEC;
    Word st[EV_MAX_STACK][VEC];
    int sp = 0;
    for (uint32_t t = 0; t < a.ntokens; ++t) {
      const uint32_t tok = a.program[t];
      if (tok < EV_MAX_OPERANDS) {
        VecIO<LB>::load(st[sp], reinterpret_cast<const Store *>(a.operands[tok]) + at);
        ++sp;
      } else if (tok == EV_COMPUTE_SHOUP) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) st[sp - 1][i] = Functor<LB, PW_COMPUTE_SHOUP>::apply(st[sp - 1][i], 0, 0, 0, p, kc);
      } else if (tok == EV_MUL_SHOUP) {
        sp -= 2;
#pragma unroll
        for (int i = 0; i < VEC; ++i) st[sp - 1][i] = Functor<LB, PW_MUL_SHOUP>::apply(st[sp - 1][i], st[sp][i], st[sp + 1][i], 0, p, kc);
      } else {
        --sp;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const Word x = st[sp - 1][i], y = st[sp][i];
          st[sp - 1][i] = tok == EV_ADD ? Functor<LB, PW_ADD>::apply(x, y, 0, 0, p, kc)
                        : tok == EV_SUB ? Functor<LB, PW_SUB>::apply(x, y, 0, 0, p, kc)
                                        : Functor<LB, PW_MUL>::apply(x, y, 0, 0, p, kc);
        }
      }
    }
    VecIO<LB>::store(dst + at, st[0]);
  }
}

template <int LB> static cudaError_t launch_eval_limb(const EvArgs &a, int num_sms, cudaStream_t stream) {
  constexpr int VEC = PW<LB>::VEC;
  const uint64_t total = (uint64_t)a.batch * (a.degree / VEC);
  if (total == 0) return cudaSuccess;
  uint64_t blocks = (total + 255) / 256;
  const uint64_t cap = (uint64_t)num_sms * 8 / a.nmoduli + 1;
  if (blocks > cap) blocks = cap;
  eval_kernel<LB><<<dim3((unsigned)blocks, a.nmoduli), 256, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_eval(int limb_bits, const EvArgs &a, int num_sms, cudaStream_t stream) {
  switch (limb_bits) {
    case 64: return launch_eval_limb<64>(a, num_sms, stream);
    case 32: return launch_eval_limb<32>(a, num_sms, stream);
    case 16: return launch_eval_limb<16>(a, num_sms, stream);
  }
  return cudaErrorInvalidValue;
}

// ---- ==  /  != on device-resident batches --------------------------------------------------------------------------
// expr::operator bool over eqmod / neqmod (ops.hpp:81-117): `a == b` is true iff ANY coefficient is equal, `a != b` iff ANY
// coefficient differs.  One CTA per polynomial pair walks the M*N limbs as 16-byte vectors; flags[b] = 1 when the
// predicate holds for polynomial b.  Reads each operand once: HBM-bound, 2*N*M*sizeof(T) bytes per polynomial.
template <int LB, bool WANT_EQUAL>
__global__ void __launch_bounds__(256) compare_kernel(const void *pa, const void *pb, uint8_t *flags, uint32_t batch, uint32_t vec_per_poly) {
  const uint4 *a = reinterpret_cast<const uint4 *>(pa), *b = reinterpret_cast<const uint4 *>(pb);
  for (uint32_t poly = blockIdx.x; poly < batch; poly += gridDim.x) {
    const size_t base = (size_t)poly * vec_per_poly;
    int hit = 0;
    for (uint32_t v = threadIdx.x; v < vec_per_poly; v += blockDim.x) {
      const uint4 x = __ldg(a + base + v), y = __ldg(b + base + v);
      const uint32_t d[4] = {x.x ^ y.x, x.y ^ y.y, x.z ^ y.z, x.w ^ y.w};
      if (LB == 64) {
        const bool e0 = (d[0] | d[1]) == 0, e1 = (d[2] | d[3]) == 0;
        hit |= WANT_EQUAL ? (e0 || e1) : (!e0 || !e1);
      } else if (LB == 32) {
#pragma unroll
        for (int i = 0; i < 4; ++i) hit |= WANT_EQUAL ? d[i] == 0 : d[i] != 0;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool e0 = (d[i] & 0xffffu) == 0, e1 = (d[i] >> 16) == 0;
          hit |= WANT_EQUAL ? (e0 || e1) : (!e0 || !e1);
        }
      }
    }
    hit = __syncthreads_or(hit);
    if (threadIdx.x == 0) flags[poly] = hit ? 1 : 0;
  }
}

cudaError_t launch_compare(int limb_bits, bool want_equal, const void *a, const void *b, uint8_t *flags, uint32_t batch, uint64_t poly_bytes,
                           int num_sms, cudaStream_t stream) {
  if (batch == 0) return cudaSuccess;
  const uint32_t vec = (uint32_t)(poly_bytes / 16);
  const unsigned grid = batch < (unsigned)num_sms * 8 ? batch : (unsigned)num_sms * 8;
#define NFLGPU_CMP(LB) \
  if (want_equal) compare_kernel<LB, true><<<grid, 256, 0, stream>>>(a, b, flags, batch, vec); \
  else compare_kernel<LB, false><<<grid, 256, 0, stream>>>(a, b, flags, batch, vec); break;
  switch (limb_bits) {
    case 64: NFLGPU_CMP(64)
    case 32: NFLGPU_CMP(32)
    case 16: NFLGPU_CMP(16)
    default: return cudaErrorInvalidValue;
  }
#undef NFLGPU_CMP
  return cudaGetLastError();
}

cudaError_t launch_pointwise(int limb_bits, int op, const PwArgs &a, int num_sms, cudaStream_t stream) {
  switch (limb_bits) {
    case 64: return launch_limb<64>(op, a, num_sms, stream);
    case 32: return launch_limb<32>(op, a, num_sms, stream);
    case 16: return launch_limb<16>(op, a, num_sms, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace nflgpu

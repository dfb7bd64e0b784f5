// Internal launcher interface of the pointwise kernels (pointwise.cu).
#ifndef NFLGPU_POINTWISE_H
#define NFLGPU_POINTWISE_H
#include <cstdint>
#include <cuda_runtime.h>

namespace nflgpu {

enum PwOp { PW_ADD = 0, PW_SUB, PW_MUL, PW_MUL_SHOUP, PW_COMPUTE_SHOUP, PW_MULADD, PW_MULADD_SHOUP };

struct PwArgs {
  void *dst;
  const void *a, *b, *c, *d;
  const uint64_t *moduli;  // [nmoduli], widened
  const uint64_t *consts;  // [nmoduli]: 64-bit limbs: Pn = low word of floor(2^128/p); 32-bit: floor(2^64/p); 16-bit: unused
  uint32_t nmoduli, degree, log2_degree, batch;
};

cudaError_t launch_pointwise(int limb_bits, int op, const PwArgs &a, int num_sms, cudaStream_t stream);

// flags[b] = (any coefficient of a[b] equal to / different from b[b])  (ops.hpp:81-117)
cudaError_t launch_compare(int limb_bits, bool want_equal, const void *a, const void *b, uint8_t *flags, uint32_t batch, uint64_t poly_bytes,
                           int num_sms, cudaStream_t stream);

// Fused evaluation of a whole expression tree in one pass (postfix program, see nflgpu_eval in include/nflgpu.h).
enum { EV_MAX_OPERANDS = 8, EV_MAX_TOKENS = 32, EV_MAX_STACK = 8 };
enum EvTok { EV_ADD = 0x10, EV_SUB = 0x11, EV_MUL = 0x12, EV_MUL_SHOUP = 0x13, EV_COMPUTE_SHOUP = 0x14 };
struct EvArgs {
  void *dst;
  const void *operands[EV_MAX_OPERANDS];
  const uint64_t *moduli, *consts;
  uint32_t nmoduli, degree, log2_degree, batch;
  uint32_t ntokens;
  uint8_t program[EV_MAX_TOKENS];
};
cudaError_t launch_eval(int limb_bits, const EvArgs &a, int num_sms, cudaStream_t stream);
// compile-time programs (eval_static.cu): true when `key` has a kernel of its own, *err = its launch status
bool launch_eval_static(int limb_bits, uint64_t key, const EvArgs &a, int num_sms, cudaStream_t stream, cudaError_t *err);

// On-device samplers (sampler.cu).
enum SampleKind { SAMPLE_UNIFORM = 0, SAMPLE_NON_UNIFORM = 1, SAMPLE_ZO = 2, SAMPLE_HWT = 3 };
struct SampleArgs {
  void *dst;
  const uint64_t *moduli;  // [nmoduli], widened
  uint32_t key[8];         // Salsa20 key, little-endian words
  uint64_t first_nonce;
  uint64_t poly_bytes, blocks_per_poly;  // blocks_per_poly = ceil(poly_bytes / 64)
  uint32_t nmoduli, log2_degree, limb_bits, batch;
  uint64_t param0, param1, param2;  // non_uniform: upper_bound, amplifier, mask;  ZO: rho;  hwt: hwt, calls per polynomial, test-only reject shift
};
// pool: scratch of the hwt sampler; hwt_used (host pointer or null): asynchronous copy of the nonces the hwt batch consumed
cudaError_t launch_sampler(int kind, const SampleArgs &a, int num_sms, cudaStream_t stream, cudaMemPool_t pool, unsigned long long *hwt_used = nullptr);

}  // namespace nflgpu
#endif

// Internal launcher interface of the CRT-lift kernels (lift.cu).
#ifndef NFLGPU_LIFT_H
#define NFLGPU_LIFT_H
#include <cstdint>
#include <cuda_runtime.h>

namespace nflgpu {

enum { LIFT_MAX_WORDS = 16, LIFT_MAX_RESIDUES = 40 };

struct LiftArgs {
  void *polys;             // limb[batch][nmoduli][degree] (mpz2poly destination; poly2mpz reads through res_ptr)
  // poly2mpz: residue cm of polynomial b starts at res_ptr[cm] + b * res_stride[cm] limbs -- one batch buffer, or slabs of
  // residue groups that may live in the memory of peer devices (read over NVLink by the lift kernel itself)
  const void *res_ptr[LIFT_MAX_RESIDUES];
  uint64_t res_stride[LIFT_MAX_RESIDUES];
  uint64_t *words;         // uint64_t[batch][degree][W], little-endian words
  const uint64_t *moduli;  // [nmoduli]
  const uint64_t *consts;  // Barrett constants (pointwise.h)
  const uint64_t *inv;     // [nmoduli]  (Q / p_cm)^-1 mod p_cm
  const uint64_t *c64;     // [nmoduli]  2^64 mod p_cm
  const uint64_t *qhat;    // [nmoduli][W]  Q / p_cm
  const uint64_t *q;       // [W]  Q = product of the moduli
  uint32_t nmoduli, log2_degree, batch;
};

// dir 0: words = lift(polys);  dir 1: polys = words mod p_cm
cudaError_t launch_lift(int limb_bits, int dir, int W, const LiftArgs &a, int num_sms, cudaStream_t stream);

}  // namespace nflgpu
#endif

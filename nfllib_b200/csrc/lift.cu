// CRT lift between the RNS representation and multi-word integers — the consumer that needs every residue of a
// polynomial on one device.
//
// Replaces, for device-resident batches, the reference's GMP code
//   poly::GMP::poly2mpz   include/nfl/gmp.hpp:183-209   x_i = sum_cm a[cm][i] * L_cm  mod Q,  Q = prod p_cm,  in [0, Q)
//   poly::GMP::mpz2poly   include/nfl/gmp.hpp:211-219   a[cm][i] = x_i mod p_cm
// A lifted coefficient is W = ceil(bits(Q)/64) little-endian 64-bit words (what mpz_export(..., -1, 8, 0, 0, x) yields).
// poly2mpz uses the other classical CRT form, x = sum_cm ((a_cm * Qhat_cm^-1 mod p_cm) * Qhat_cm) - k*Q with
// Qhat_cm = Q / p_cm: every term is below Q, so k < nmoduli and the result is the same unique representative in [0, Q)
// that the reference's Shoup-style big-number reduction produces.  One thread per coefficient, integer ALU only.
#include "lift.h"
#include "modmul.cuh"

namespace nflgpu {

template <int LB> struct LimbIO;
template <> struct LimbIO<64> { typedef uint64_t T; };
template <> struct LimbIO<32> { typedef uint32_t T; };
template <> struct LimbIO<16> { typedef uint16_t T; };

template <int LB, int W>
__global__ void __launch_bounds__(128) poly2words_kernel(const LiftArgs a) {
  typedef typename LimbIO<LB>::T Store;
  typedef typename PW<LB>::Word Word;
  const uint64_t total = (uint64_t)a.batch << a.log2_degree;
  const uint64_t degree = 1ull << a.log2_degree;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t b = t >> a.log2_degree, i = t & (degree - 1);
    uint64_t acc[W + 1];
#pragma unroll
    for (int k = 0; k <= W; ++k) acc[k] = 0;
    for (uint32_t cm = 0; cm < a.nmoduli; ++cm) {
      const Word p = (Word)a.moduli[cm];
      const Word limb = (Word)__ldg(reinterpret_cast<const Store *>(a.res_ptr[cm]) + b * a.res_stride[cm] + i);
      const uint64_t v = PW<LB>::mulmod(limb, (Word)a.inv[cm], p, a.consts[cm]);  // a_cm * Qhat_cm^-1 mod p_cm
      const uint64_t *qh = a.qhat + (uint64_t)cm * W;
      uint64_t carry = 0;
#pragma unroll
      for (int k = 0; k < W; ++k) {  // acc += v * Qhat_cm
        const uint64_t lo = v * qh[k], hi = __umul64hi(v, qh[k]);
        uint64_t s = acc[k] + lo;
        uint64_t c1 = s < lo;
        s += carry;
        c1 += s < carry;
        acc[k] = s;
        carry = hi + c1;
      }
      acc[W] += carry;
    }
    // acc < nmoduli * Q: subtract Q until it fits (at most nmoduli - 1 times)
    for (uint32_t it = 0; it < a.nmoduli; ++it) {
      bool ge = acc[W] != 0;
      if (!ge) {
        ge = true;
#pragma unroll
        for (int k = W - 1; k >= 0; --k) {
          if (acc[k] != a.q[k]) { ge = acc[k] > a.q[k]; break; }
        }
      }
      if (!ge) break;
      uint64_t borrow = 0;
#pragma unroll
      for (int k = 0; k < W; ++k) {
        const uint64_t d = acc[k] - a.q[k], b1 = acc[k] < a.q[k];
        const uint64_t d2 = d - borrow, b2 = d < borrow;
        acc[k] = d2;
        borrow = b1 + b2;
      }
      acc[W] -= borrow;
    }
    uint64_t *dst = a.words + t * W;
#pragma unroll
    for (int k = 0; k < W; ++k) dst[k] = acc[k];
  }
}

template <int LB, int W>
__global__ void __launch_bounds__(128) words2poly_kernel(const LiftArgs a) {
  typedef typename LimbIO<LB>::T Store;
  const uint64_t total = (uint64_t)a.batch << a.log2_degree;
  const uint64_t degree = 1ull << a.log2_degree;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t b = t >> a.log2_degree, i = t & (degree - 1);
    uint64_t w[W];
#pragma unroll
    for (int k = 0; k < W; ++k) w[k] = a.words[t * W + k];
    Store *dst = reinterpret_cast<Store *>(a.polys) + b * a.nmoduli * degree + i;
    for (uint32_t cm = 0; cm < a.nmoduli; ++cm) {
      const uint64_t p = a.moduli[cm], c64 = a.c64[cm];  // c64 = 2^64 mod p
      uint64_t r = 0;
#pragma unroll
      for (int k = W - 1; k >= 0; --k) {  // Horner in base 2^64
        uint64_t d = w[k];
        if (LB == 64) {
          while (d >= p) d -= p;  // d < 2^64 < 5p: a few subtractions
          r = PW<64>::mulmod(r, c64, p, a.consts[cm]) + d;
          r = r >= p ? r - p : r;
        } else {
          r = (r * c64 + d % p) % p;  // r, c64 < 2^30: the products fit 64 bits
        }
      }
      dst[(uint64_t)cm * degree] = (Store)r;
    }
  }
}

template <int LB, int W> static cudaError_t launch_lift_w(int dir, const LiftArgs &a, int num_sms, cudaStream_t stream) {
  const uint64_t total = (uint64_t)a.batch << a.log2_degree;
  if (total == 0) return cudaSuccess;
  uint64_t blocks = (total + 127) / 128;
  if (blocks > (uint64_t)num_sms * 16) blocks = (uint64_t)num_sms * 16;
  if (dir == 0) poly2words_kernel<LB, W><<<(unsigned)blocks, 128, 0, stream>>>(a);
  else words2poly_kernel<LB, W><<<(unsigned)blocks, 128, 0, stream>>>(a);
  return cudaGetLastError();
}

template <int LB> static cudaError_t launch_lift_limb(int dir, int W, const LiftArgs &a, int num_sms, cudaStream_t stream) {
  switch (W) {
#define NFLGPU_LIFT_CASE(K) case K: return launch_lift_w<LB, K>(dir, a, num_sms, stream);
    NFLGPU_LIFT_CASE(1) NFLGPU_LIFT_CASE(2) NFLGPU_LIFT_CASE(3) NFLGPU_LIFT_CASE(4) NFLGPU_LIFT_CASE(5) NFLGPU_LIFT_CASE(6)
    NFLGPU_LIFT_CASE(7) NFLGPU_LIFT_CASE(8) NFLGPU_LIFT_CASE(9) NFLGPU_LIFT_CASE(10) NFLGPU_LIFT_CASE(11) NFLGPU_LIFT_CASE(12)
    NFLGPU_LIFT_CASE(13) NFLGPU_LIFT_CASE(14) NFLGPU_LIFT_CASE(15) NFLGPU_LIFT_CASE(16)
#undef NFLGPU_LIFT_CASE
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_lift(int limb_bits, int dir, int W, const LiftArgs &a, int num_sms, cudaStream_t stream) {
  switch (limb_bits) {
    case 64: return launch_lift_limb<64>(dir, W, a, num_sms, stream);
    case 32: return launch_lift_limb<32>(dir, W, a, num_sms, stream);
    case 16: return launch_lift_limb<16>(dir, W, a, num_sms, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace nflgpu

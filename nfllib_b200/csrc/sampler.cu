// On-device uniform sampler: poly::set(nfl::uniform) for a whole batch, born in HBM.
//
// Replaces, for device-resident batches, the reference's
//   poly::set(uniform const&)              core.hpp:150-187   (mask every limb to the modulus' bit length, one conditional subtract)
//   nfl::fastrandombytes                    lib/prng/fastrandombytes.cpp:21-34  (Salsa20 keystream, one 64-bit nonce per call)
//   nfl_crypto_stream_salsa20_amd64_xmm6    lib/prng/*.s       (Salsa20/20, D. J. Bernstein's public specification)
// Polynomial i of the batch is filled from the keystream (key, first_nonce + i), exactly what `batch` successive
// poly::set(uniform) calls produce, so that with the same key the device batch is bit-identical to the reference's
// draws (tests compare against the reference itself run with a fixed key).  One thread = one 64-byte Salsa20 block =
// 8 / 16 / 32 limbs, written with four 16-byte stores; integer ALU only (add / rotate / xor).
#include "pointwise.h"

namespace nflgpu {

__device__ __forceinline__ uint32_t rotl32(uint32_t v, int c) { return __funnelshift_l(v, v, c); }

__global__ void __launch_bounds__(256) uniform_kernel(const SampleArgs a) {
  const uint64_t total = (uint64_t)a.batch * a.blocks_per_poly;
  for (uint64_t gb = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; gb < total; gb += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t poly = gb / a.blocks_per_poly, blk = gb - poly * a.blocks_per_poly;
    const uint64_t nonce = a.first_nonce + poly;
    uint32_t in[16], x[16];
    in[0] = 0x61707865u; in[5] = 0x3320646eu; in[10] = 0x79622d32u; in[15] = 0x6b206574u;  // "expand 32-byte k"
    in[1] = a.key[0]; in[2] = a.key[1]; in[3] = a.key[2]; in[4] = a.key[3];
    in[11] = a.key[4]; in[12] = a.key[5]; in[13] = a.key[6]; in[14] = a.key[7];
    in[6] = (uint32_t)nonce; in[7] = (uint32_t)(nonce >> 32); in[8] = (uint32_t)blk; in[9] = (uint32_t)(blk >> 32);
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = in[i];
#define NFLGPU_QR(A, B, C, D) \
  x[B] ^= rotl32(x[A] + x[D], 7); x[C] ^= rotl32(x[B] + x[A], 9); x[D] ^= rotl32(x[C] + x[B], 13); x[A] ^= rotl32(x[D] + x[C], 18);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      NFLGPU_QR(0, 4, 8, 12) NFLGPU_QR(5, 9, 13, 1) NFLGPU_QR(10, 14, 2, 6) NFLGPU_QR(15, 3, 7, 11)
      NFLGPU_QR(0, 1, 2, 3) NFLGPU_QR(5, 6, 7, 4) NFLGPU_QR(10, 11, 8, 9) NFLGPU_QR(15, 12, 13, 14)
    }
#undef NFLGPU_QR
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] += in[i];

    // mask + conditional subtract per limb (core.hpp:163-176); all limbs of a 64-byte block belong to one residue
    // whenever degree * limb_bytes >= 64, otherwise look the residue up per limb
    const uint64_t byte0 = blk * 64;
    unsigned char *dst = reinterpret_cast<unsigned char *>(a.dst) + poly * a.poly_bytes + byte0;
    if (a.limb_bits == 64) {
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        uint64_t w[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint64_t limb = (byte0 >> 3) + 2 * v + h;
          const uint64_t p = a.moduli[(limb >> a.log2_degree) % a.nmoduli];
          const uint64_t mask = (2ull << (63 - __clzll(p))) - 1;
          uint64_t t = (((uint64_t)x[4 * v + 2 * h + 1] << 32) | x[4 * v + 2 * h]) & mask;
          w[h] = t >= p ? t - p : t;
        }
        if (byte0 + 16 * v < a.poly_bytes) *reinterpret_cast<ulonglong2 *>(dst + 16 * v) = make_ulonglong2(w[0], w[1]);
      }
    } else if (a.limb_bits == 32) {
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        uint32_t w[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const uint64_t limb = (byte0 >> 2) + 4 * v + h;
          const uint32_t p = (uint32_t)a.moduli[(limb >> a.log2_degree) % a.nmoduli];
          const uint32_t mask = (2u << (31 - __clz(p))) - 1;
          const uint32_t t = x[4 * v + h] & mask;
          w[h] = t >= p ? t - p : t;
        }
        if (byte0 + 16 * v < a.poly_bytes) *reinterpret_cast<uint4 *>(dst + 16 * v) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    } else {
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        uint32_t w[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          uint32_t packed = 0;
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const uint64_t limb = (byte0 >> 1) + 8 * v + 2 * h + s;
            const uint32_t p = (uint32_t)a.moduli[(limb >> a.log2_degree) % a.nmoduli];
            const uint32_t mask = (2u << (31 - __clz(p))) - 1;
            uint32_t t = ((x[4 * v + h] >> (16 * s)) & 0xffffu) & mask;
            t = t >= p ? t - p : t;
            packed |= t << (16 * s);
          }
          w[h] = packed;
        }
        if (byte0 + 16 * v < a.poly_bytes) *reinterpret_cast<uint4 *>(dst + 16 * v) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
}

cudaError_t launch_uniform(const SampleArgs &a, int num_sms, cudaStream_t stream) {
  const uint64_t total = (uint64_t)a.batch * a.blocks_per_poly;
  if (total == 0) return cudaSuccess;
  uint64_t blocks = (total + 255) / 256;
  if (blocks > (uint64_t)num_sms * 16) blocks = (uint64_t)num_sms * 16;
  uniform_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace nflgpu

// On-device samplers: poly::set(nfl::uniform / non_uniform / ZO_dist) for a whole batch, born in HBM.
//
// Replaces, for device-resident batches, the reference's
//   poly::set(uniform const&)              core.hpp:150-187   (mask every limb to the modulus' bit length, one conditional subtract)
//   poly::set(non_uniform const&)          core.hpp:190-278   (centred bounded noise, same value in every residue)
//   poly::set(ZO_dist const&)              core.hpp:338-349   ({-1,0,1} from one keystream byte per coefficient)
//   poly::set(hwt_dist const&)             core.hpp:355-392   (exactly hwt coefficients +-1, reservoir sampling)
//   nfl::fastrandombytes                    lib/prng/fastrandombytes.cpp:21-34  (Salsa20 keystream, one 64-bit nonce per call)
//   nfl_crypto_stream_salsa20_amd64_xmm6    lib/prng/*.s       (Salsa20/20, D. J. Bernstein's public specification)
// Polynomial i of the batch is filled from the keystream (key, first_nonce + i), exactly what `batch` successive
// poly::set(uniform) calls produce, so that with the same key the device batch is bit-identical to the reference's
// draws (tests compare against the reference itself run with a fixed key).  One thread = one 64-byte Salsa20 block =
// 8 / 16 / 32 limbs, written with four 16-byte stores; integer ALU only (add / rotate / xor).
#include "gaussian.h"
#include "pointwise.h"

namespace nflgpu {

__device__ __forceinline__ uint32_t rotl32(uint32_t v, int c) { return __funnelshift_l(v, v, c); }

// one 64-byte Salsa20/20 keystream block (key, nonce, block counter) -> x[16] little-endian words
__device__ __forceinline__ void salsa20_block(const uint32_t (&key)[8], uint64_t nonce, uint64_t blk, uint32_t (&x)[16]) {
  uint32_t in[16];
  in[0] = 0x61707865u; in[5] = 0x3320646eu; in[10] = 0x79622d32u; in[15] = 0x6b206574u;  // "expand 32-byte k"
  in[1] = key[0]; in[2] = key[1]; in[3] = key[2]; in[4] = key[3];
  in[11] = key[4]; in[12] = key[5]; in[13] = key[6]; in[14] = key[7];
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
}

template <class T> __device__ __forceinline__ void store_limb(void *base, uint64_t index, uint64_t v) { reinterpret_cast<T *>(base)[index] = (T)v; }
__device__ __forceinline__ void store_any(void *base, uint32_t limb_bits, uint64_t index, uint64_t v) {
  if (limb_bits == 64) store_limb<uint64_t>(base, index, v);
  else if (limb_bits == 32) store_limb<uint32_t>(base, index, v);
  else store_limb<uint16_t>(base, index, v);
}

// poly::set(non_uniform) core.hpp:190-278: one keystream of degree limbs per polynomial; coefficient i is the same
// centred bounded value in every residue (p_cm - |v| for the negative half), optionally amplified.
__global__ void __launch_bounds__(256) non_uniform_kernel(const SampleArgs a) {
  const uint32_t limb_bytes = a.limb_bits / 8, per_block = 64 / limb_bytes;
  const uint64_t degree = 1ull << a.log2_degree;
  const uint64_t total = (uint64_t)a.batch * a.blocks_per_poly;
  const uint64_t wrap = a.limb_bits == 64 ? ~0ull : ((1ull << a.limb_bits) - 1);
  const uint64_t two_ub_m1 = 2 * a.param0 - 1;
  for (uint64_t gb = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; gb < total; gb += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t poly = gb / a.blocks_per_poly, blk = gb - poly * a.blocks_per_poly;
    uint32_t x[16];
    salsa20_block(a.key, a.first_nonce + poly, blk, x);
    for (uint32_t t = 0; t < per_block; ++t) {
      const uint64_t i = blk * per_block + t;
      if (i >= degree) break;
      uint64_t raw;
      if (a.limb_bits == 64) raw = ((uint64_t)x[2 * t + 1] << 32) | x[2 * t];
      else if (a.limb_bits == 32) raw = x[t];
      else raw = (x[t >> 1] >> (16 * (t & 1))) & 0xffffu;
      uint64_t tmp = raw & a.param2;                                   // core.hpp:218-219,226
      if (tmp >= two_ub_m1) tmp -= two_ub_m1;                           // core.hpp:232-234
      for (uint32_t cm = 0; cm < a.nmoduli; ++cm) {
        const uint64_t v = tmp >= a.param0 ? a.moduli[cm] + tmp * a.param1 - two_ub_m1 * a.param1 : tmp * a.param1;  // core.hpp:236-246,262-271
        store_any(reinterpret_cast<unsigned char *>(a.dst) + poly * a.poly_bytes, a.limb_bits, (uint64_t)cm * degree + i, v & wrap);
      }
    }
  }
}

// poly::set(ZO_dist) core.hpp:338-349: one keystream BYTE per coefficient; {-1, 0, +1} stored as p-1 / 0 / p+1
// (the reference really stores p + 1 for +1, core.hpp:347).
__global__ void __launch_bounds__(256) zo_kernel(const SampleArgs a) {
  const uint64_t degree = 1ull << a.log2_degree;
  const uint64_t total = (uint64_t)a.batch * a.blocks_per_poly;
  for (uint64_t gb = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; gb < total; gb += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t poly = gb / a.blocks_per_poly, blk = gb - poly * a.blocks_per_poly;
    uint32_t x[16];
    salsa20_block(a.key, a.first_nonce + poly, blk, x);
    for (uint32_t t = 0; t < 64; ++t) {
      const uint64_t i = blk * 64 + t;
      if (i >= degree) break;
      const uint32_t r = (x[t >> 2] >> (8 * (t & 3))) & 0xffu;
      for (uint32_t cm = 0; cm < a.nmoduli; ++cm) {
        const uint64_t v = r <= a.param0 ? (a.moduli[cm] - 1) + (r & 2) : 0;
        store_any(reinterpret_cast<unsigned char *>(a.dst) + poly * a.poly_bytes, a.limb_bits, (uint64_t)cm * degree + i, v);
      }
    }
  }
}

// poly::set(hwt_dist) core.hpp:355-392: reservoir sampling is sequential inside a polynomial, so one thread draws one
// polynomial (the batch supplies the parallelism).  `hit` and `bitmap` are scratch: the reservoir slots and the set of
// selected positions (the reference sorts the slots; scanning the bitmap in ascending order pairs position and sign word
// identically).  A polynomial normally consumes param1 = ceil((N - hwt) / hwt) + 1 fastrandombytes calls (nonces); an index
// rejected by the rejection sampling (probability < 2^-44 per draw) can push it past a refill boundary and cost one more, which
// shifts the start nonce of every later polynomial of the reference's sequential stream.  hwt_one returns the calls actually
// made; hwt_kernel assumes the normal count for the starts, hwt_repair_kernel re-draws whatever a longer polynomial displaced.
// param2 (testing aid, 0 in production) cuts the top 2^-param2 part off the acceptance range so that tests can force the case.
__device__ uint64_t hwt_one(const SampleArgs &a, uint64_t poly, uint64_t start_nonce, uint32_t *hit, uint32_t *bitmap) {
  const uint64_t degree = 1ull << a.log2_degree;
  const uint32_t hwt = (uint32_t)a.param0, shift = (uint32_t)a.param2;
  uint64_t nonce = start_nonce, cur = 0, blk = 0;
  uint32_t x[16], avail = 0, widx = 8;
  for (uint32_t k = 0; k < hwt; ++k) hit[k] = k;
  for (uint64_t k = hwt; k < degree; ++k) {
    const uint64_t reject = 0xffffffffffffffffull / k;
    uint64_t pos;
    for (;;) {
      if (avail == 0) { avail = hwt; cur = nonce++; blk = 0; widx = 8; }
      if (widx == 8) { salsa20_block(a.key, cur, blk++, x); widx = 0; }
      pos = ((uint64_t)x[2 * widx + 1] << 32) | x[2 * widx];
      ++widx; --avail;
      if (pos <= reject * k - (shift ? (reject * k) >> shift : 0)) { pos %= k; break; }
    }
    if (pos < hwt) hit[pos] = (uint32_t)k;
  }
  for (uint32_t k = 0; k < hwt; ++k) atomicOr(&bitmap[hit[k] >> 5], 1u << (hit[k] & 31));
  cur = nonce++; blk = 0; widx = 8;  // the sign keystream: one more fastrandombytes call
  unsigned char *dst = reinterpret_cast<unsigned char *>(a.dst) + poly * a.poly_bytes;  // zeroed by the caller
  for (uint64_t wd = 0; wd < (degree + 31) / 32; ++wd) {
    uint32_t bits = bitmap[wd];
    while (bits) {
      const uint32_t bit = __ffs(bits) - 1;
      bits &= bits - 1;
      if (widx == 8) { salsa20_block(a.key, cur, blk++, x); widx = 0; }
      const uint32_t sign = x[2 * widx] & 2u;
      ++widx;
      for (uint32_t cm = 0; cm < a.nmoduli; ++cm)
        store_any(dst, a.limb_bits, (uint64_t)cm * degree + wd * 32 + bit, (a.moduli[cm] - 1) + sign);
    }
  }
  return nonce - start_nonce;
}

__global__ void __launch_bounds__(64) hwt_kernel(const SampleArgs a, uint32_t *hit_all, uint32_t *bitmap_all, uint32_t *used) {
  const uint64_t degree = 1ull << a.log2_degree;
  const uint64_t poly = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (poly >= a.batch) return;
  // bitmap and destination were zeroed by the launcher
  used[poly] = (uint32_t)hwt_one(a, poly, a.first_nonce + poly * a.param1, hit_all + poly * a.param0, bitmap_all + poly * ((degree + 31) / 32));
}

// One CTA.  Invariant: the polynomials [done, batch) have been drawn with start nonces start + (i - done) * param1.  Find the
// first of them that used another number of nonces; everything up to and including it is final, the stream continues right
// after it, and the polynomials behind it are drawn again from there.  Ends when no polynomial deviates (at once, normally).
// result[0] = nonces the whole batch consumed.
__global__ void __launch_bounds__(256) hwt_repair_kernel(const SampleArgs a, uint32_t *hit_all, uint32_t *bitmap_all, uint32_t *used,
                                                         unsigned long long *result) {
  __shared__ unsigned int first_bad;
  const uint64_t degree = 1ull << a.log2_degree, words = (degree + 31) / 32;
  uint64_t done = 0, start = a.first_nonce;
  for (;;) {
    if (threadIdx.x == 0) first_bad = 0xffffffffu;
    __syncthreads();
    for (uint64_t i = done + threadIdx.x; i < a.batch; i += blockDim.x)
      if (used[i] != (uint32_t)a.param1) { atomicMin(&first_bad, (unsigned int)i); break; }
    __syncthreads();
    const uint64_t bad = first_bad;
    __syncthreads();
    if (bad == 0xffffffffu) {
      if (threadIdx.x == 0) result[0] = start + (a.batch - done) * a.param1 - a.first_nonce;
      return;
    }
    start += (bad - done) * a.param1 + used[bad];
    done = bad + 1;
    for (uint64_t i = done + threadIdx.x; i < a.batch; i += blockDim.x) {
      uint32_t *bitmap = bitmap_all + i * words;
      for (uint64_t w = 0; w < words; ++w) bitmap[w] = 0;
      unsigned char *dst = reinterpret_cast<unsigned char *>(a.dst) + i * a.poly_bytes;
      for (uint64_t b = 0; b < a.poly_bytes; b += 4) *reinterpret_cast<uint32_t *>(dst + b) = 0;  // core.hpp:383
      used[i] = (uint32_t)hwt_one(a, i, start + (i - done) * a.param1, hit_all + i * a.param0, bitmap);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) uniform_kernel(const SampleArgs a) {
  const uint64_t total = (uint64_t)a.batch * a.blocks_per_poly;
  for (uint64_t gb = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; gb < total; gb += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t poly = gb / a.blocks_per_poly, blk = gb - poly * a.blocks_per_poly;
    uint32_t x[16];
    salsa20_block(a.key, a.first_nonce + poly, blk, x);

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

// ---- poly::set(gaussian) core.hpp:291-325 over FastGaussianNoise::getNoise (prng/FastGaussianNoise.hpp:478-613) ----------
//
// getNoise() is sequential twice over: inside a polynomial every output consumes 1, 2 or word_precision look-up words of one
// keystream, so where an output starts depends on all outputs before it; and across polynomials the NONCE a draw starts with
// depends on how many times the draws before it refilled their buffer (refills are data dependent, :601-610).  Both chains
// are turned into "evaluate every possible start in parallel, then follow the chain":
//   1. gauss_positions: for every nonce n of a window and EVERY position j of its keystream buffer, the output an evaluation
//      starting at j would produce and the words it would consume (one CTA per nonce: Salsa20 blocks into shared memory,
//      then one look-up evaluation per thread and position);
//   2. gauss_walk: thread c follows position -> position + consumed for the draw that would start at nonce first_nonce + c
//      (the consumed-words rows of its nonces are staged in shared memory, so a step is one shared-memory load and one
//      fire-and-forget store: no global load sits on the chain), records WHERE each of the draw's outputs was evaluated
//      and the draw's number of fastrandombytes calls;
//   3. gauss_chain: one thread follows nonce -> nonce + calls(nonce) from first_nonce: the candidates the reference's
//      sequential draws really are (calls staged in shared memory);
//   4. gauss_expand: gathers the outputs of the chosen candidates into the batch, amplified, negative values stored as
//      p + v in every residue (coalesced stores).
__global__ void __launch_bounds__(128) gauss_positions_kernel(const GaussArgs a) {
  extern __shared__ __align__(16) unsigned char gsm[];
  uint32_t *ks = reinterpret_cast<uint32_t *>(gsm);  // keystream of this nonce, whole 64-byte blocks
  const uint32_t row = blockIdx.x;                   // nonce first_nonce + row
  const uint32_t words = (uint32_t)a.words_per_fill, wp = a.wp, depth = a.depth, ib = a.in_bytes;
  const uint32_t nblk = (words * ib + 63) / 64;
  for (uint32_t blk = threadIdx.x; blk < nblk; blk += blockDim.x) {
    uint32_t x[16];
    salsa20_block(a.key, a.first_nonce + row, blk, x);
#pragma unroll
    for (int i = 0; i < 16; ++i) ks[blk * 16 + i] = x[i];
  }
  __syncthreads();
  auto word = [&](uint32_t j) -> uint32_t {
    const uint32_t byte = j * ib, w = ks[byte >> 2] >> (8 * (byte & 3));
    return ib == 1 ? (w & 0xffu) : (w & 0xffffu);  // 16-bit words never straddle a 32-bit word
  };
  int32_t *val = a.pos_val + (uint64_t)row * words;
  uint8_t *adv = a.pos_adv + (uint64_t)row * a.walk_stride;
  for (uint32_t j = threadIdx.x; j < words; j += blockDim.x) {
    int32_t output = 0;
    uint32_t consumed = 1;
    if (j + wp < words) {  // getNoise() only ever starts an output where a full-precision comparison still fits (:601)
      GaussLutEntry e = a.lut[word(j)];
      const bool flagged1 = e.sub >= 0;
      if (flagged1 && depth == 2) { e = a.lut[(uint64_t)e.sub * a.lu_size + word(j + 1)]; consumed = 2; }
      output = e.val;
      if (e.sub >= 0) {  // walk the barriers of this entry: +1 for every barrier not above the noise (cmp(), :617-628)
        for (uint32_t b = e.bstart; b < e.bstart + e.bcount; ++b) {
          const unsigned char *brow = a.barriers + (uint64_t)b * wp * ib;
          int cmp = 0;
          for (uint32_t i = 0; i < wp && cmp == 0; ++i) {
            const uint32_t bw = ib == 1 ? brow[i] : (uint32_t)brow[2 * i] | ((uint32_t)brow[2 * i + 1] << 8), nw = word(j + i);
            cmp = bw > nw ? 1 : (bw < nw ? -1 : 0);
          }
          if (cmp == 1) break;
          ++output;
        }
        consumed = wp;
      }
    }
    val[j] = output;
    adv[j] = (uint8_t)consumed;
  }
}

__global__ void __launch_bounds__(GAUSS_WALK_THREADS) gauss_walk_kernel(const GaussArgs a) {
  extern __shared__ __align__(16) unsigned char gsm[];
  const uint32_t words = (uint32_t)a.words_per_fill, wp = a.wp, stride = a.walk_stride, rows = a.rows, window = a.window;
  const uint32_t c0 = blockIdx.x * blockDim.x, staged = a.walk_rows;  // rows c0 .. c0 + staged - 1 live in shared memory
  const uint8_t *__restrict__ adv = a.pos_adv;
  {  // rows have the same 16-byte-multiple pitch in global and in shared memory: one flat vector copy
    const uint32_t nrows = c0 + staged <= rows ? staged : (c0 < rows ? rows - c0 : 0);
    const uint4 *src = reinterpret_cast<const uint4 *>(adv + (uint64_t)c0 * stride);
    uint4 *dst = reinterpret_cast<uint4 *>(gsm);
    for (uint32_t i = threadIdx.x; i < nrows * (stride / 16); i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const uint32_t c = c0 + threadIdx.x;
  if (c >= window) return;
  const uint32_t degree = 1u << a.log2_degree;
  // output k of candidate c goes to cand_idx[k][c]: the 32 candidates of a warp write one 128-byte line per step
  uint32_t *__restrict__ out = a.cand_idx + c;
  uint32_t n = c, pos = 0, calls = 1;
  const unsigned char *row = gsm + (size_t)threadIdx.x * stride;  // row of nonce n while it is staged
  bool in_smem = threadIdx.x < staged;  // large degrees stage fewer rows than the CTA has candidates: the others read global memory
  for (uint32_t k = 0; k < degree; ++k) {
    if (n >= rows) { calls = 0x40000000u; break; }  // ran past the evaluated nonces: the chain treats it as "window too small"
    out[(uint64_t)k * window] = n * words + pos;      // index into pos_val (rows * words < 2^32: the caller's chunking bounds it)
    pos += in_smem ? row[pos] : __ldg(adv + (uint64_t)n * stride + pos);
    if (pos + wp >= words) {                          // :601-610: a fresh buffer from the next nonce
      pos = 0; ++n; ++calls;
      in_smem = n - c0 < staged;
      row += stride;
    }
  }
  a.cand_calls[c] = calls;
}

// The nonce chain by pointer jumping: J_k[i] = where the chain is 2^k draws after nonce index i (index `window` absorbs
// everything that leaves the window).  With chosen[0 .. 2^k) known, chosen[b + 2^k] = J_k[chosen[b]]; then J_{k+1} = J_k o J_k.
// log2(batch) + 1 rounds of fully parallel shared-memory work instead of batch dependent steps.
__global__ void __launch_bounds__(1024) gauss_chain_kernel(const GaussArgs a) {
  extern __shared__ __align__(16) unsigned char gsm[];
  const uint32_t W = a.window;
  uint32_t *J = reinterpret_cast<uint32_t *>(gsm), *K = J + (W + 1);
  for (uint32_t i = threadIdx.x; i <= W; i += blockDim.x) {
    uint64_t v = W;
    if (i < W) { v = (uint64_t)i + a.cand_calls[i]; if (v > W) v = W; }
    J[i] = (uint32_t)v;
  }
  if (threadIdx.x == 0) a.chosen[0] = 0;
  __syncthreads();
  for (uint32_t len = 1; len < a.batch; len <<= 1) {
    for (uint32_t b = threadIdx.x; b < len && b + len < a.batch; b += blockDim.x) a.chosen[b + len] = J[a.chosen[b]];
    for (uint32_t i = threadIdx.x; i <= W; i += blockDim.x) K[i] = J[J[i]];
    __syncthreads();
    uint32_t *t = J; J = K; K = t;
  }
  if (threadIdx.x == 0) {
    const uint32_t last = a.chosen[a.batch - 1];
    const uint32_t calls = last < W ? a.cand_calls[last] : 0x40000000u;
    a.result[0] = (uint64_t)last + calls;
    a.result[1] = calls >= 0x40000000u ? 1 : 0;
  }
}

// One CTA = 32 polynomials x 32 coefficients: the evaluation indices are read along the polynomial axis (cand_idx[k][c] with
// c = chosen[poly], nearly consecutive), the values are gathered from pos_val, and the tile is written along the coefficient
// axis (32 consecutive limbs of one residue row per warp store).
__global__ void __launch_bounds__(1024) gauss_expand_kernel(const GaussArgs a) {
  __shared__ int32_t tile[32][33];
  const uint32_t degree = 1u << a.log2_degree, window = a.window;
  const uint64_t wrap = a.limb_bits == 64 ? ~0ull : ((1ull << a.limb_bits) - 1);
  const uint32_t tiles_i = degree / 32 ? degree / 32 : 1, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const uint32_t ntiles = ((a.batch + 31) / 32) * tiles_i;
  for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const uint32_t p0 = (t / tiles_i) * 32, i0 = (t % tiles_i) * 32;
    {  // tx runs over polynomials, ty over coefficients
      const uint32_t poly = p0 + tx, i = i0 + ty;
      int32_t v = 0;
      if (poly < a.batch && i < degree) {
        // (a draw that ran out of the evaluated nonces leaves its remaining indices unwritten; the chain reports that and the
        //  host retries, but this kernel has been launched already: keep whatever it reads inside pos_val)
        const uint32_t idx = a.cand_idx[(uint64_t)i * window + min(a.chosen[poly], window - 1)];
        v = a.pos_val[min(idx, a.rows * (uint32_t)a.words_per_fill - 1u)];
      }
      tile[ty][tx] = v;
    }
    __syncthreads();
    {  // tx runs over coefficients, ty over polynomials
      const uint32_t poly = p0 + ty, i = i0 + tx;
      if (poly < a.batch && i < degree) {
        // rnd[i] is signed_value_type: getNoise stores the low limb_bits of the output, rnd *= amplifier wraps at the limb width
        uint64_t v = (uint64_t)(int64_t)tile[tx][ty] & wrap;
        if (a.amplifier != 1) v = (v * a.amplifier) & wrap;
        const bool neg = (v >> (a.limb_bits - 1)) & 1;
        for (uint32_t cm = 0; cm < a.nmoduli; ++cm)
          store_any(reinterpret_cast<unsigned char *>(a.dst) + (uint64_t)poly * a.poly_bytes, a.limb_bits, (uint64_t)cm * degree + i,
                    (neg ? a.moduli[cm] + v : v) & wrap);
      }
    }
    __syncthreads();
  }
}

cudaError_t launch_gaussian(GaussArgs a, int device, int num_sms, cudaStream_t stream) {
  if (a.batch == 0) return cudaSuccess;
  const uint32_t words = (uint32_t)a.words_per_fill;
  static bool attr_set_dev[64] = {false};  // per device; racing first calls set the same values
  if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
  bool &attr_set = attr_set_dev[device];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gauss_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GAUSS_SMEM_BUDGET);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gauss_positions_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GAUSS_SMEM_BUDGET);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gauss_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GAUSS_SMEM_BUDGET);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const size_t ks_bytes = (size_t)((words * a.in_bytes + 63) / 64) * 64;
  if (ks_bytes > GAUSS_SMEM_BUDGET) return cudaErrorInvalidValue;  // the launcher's caller checks this bound first
  a.walk_stride = (words + 15) & ~15u;  // pitch of a pos_adv row, in global and in shared memory
  gauss_positions_kernel<<<a.rows, 128, ks_bytes, stream>>>(a);
  // rows of consumed-words staged per walk CTA: its own candidates' nonces plus two for refills, as many as fit
  uint32_t staged = GAUSS_WALK_THREADS + 2;
  if ((size_t)staged * a.walk_stride > GAUSS_SMEM_BUDGET) staged = (uint32_t)(GAUSS_SMEM_BUDGET / a.walk_stride);
  a.walk_rows = staged;
  gauss_walk_kernel<<<(a.window + GAUSS_WALK_THREADS - 1) / GAUSS_WALK_THREADS, GAUSS_WALK_THREADS, (size_t)staged * a.walk_stride, stream>>>(a);
  if ((size_t)(a.window + 1) * 8 > GAUSS_SMEM_BUDGET) return cudaErrorInvalidValue;  // the caller's chunking bounds the window
  gauss_chain_kernel<<<1, 1024, (size_t)(a.window + 1) * 8, stream>>>(a);
  const uint64_t degree = 1ull << a.log2_degree;
  uint64_t tiles = (uint64_t)((a.batch + 31) / 32) * (degree / 32 ? degree / 32 : 1);
  if (tiles > (uint64_t)num_sms * 8) tiles = (uint64_t)num_sms * 8;
  gauss_expand_kernel<<<(unsigned)tiles, 1024, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_hwt(const SampleArgs &a, cudaStream_t stream, cudaMemPool_t pool, unsigned long long *result) {
  if (a.batch == 0) return cudaSuccess;
  const uint64_t degree = 1ull << a.log2_degree, hwt = a.param0;
  const size_t hit_bytes = ((size_t)a.batch * hwt * 4 + 15) & ~(size_t)15, bm_bytes = ((size_t)a.batch * ((degree + 31) / 32) * 4 + 15) & ~(size_t)15;
  const size_t used_bytes = ((size_t)a.batch * 4 + 15) & ~(size_t)15;  // (every part a multiple of 16 bytes: the 8-byte result stays aligned)
  unsigned char *scratch = nullptr;
  cudaError_t e = cudaMallocFromPoolAsync(reinterpret_cast<void **>(&scratch), hit_bytes + bm_bytes + used_bytes + 16, pool, stream);
  if (e != cudaSuccess) return e;
  uint32_t *hit = reinterpret_cast<uint32_t *>(scratch), *bitmap = reinterpret_cast<uint32_t *>(scratch + hit_bytes);
  uint32_t *used = reinterpret_cast<uint32_t *>(scratch + hit_bytes + bm_bytes);
  unsigned long long *res = reinterpret_cast<unsigned long long *>(scratch + hit_bytes + bm_bytes + used_bytes);
  if ((e = cudaMemsetAsync(bitmap, 0, bm_bytes, stream)) == cudaSuccess &&
      (e = cudaMemsetAsync(a.dst, 0, (size_t)a.batch * a.poly_bytes, stream)) == cudaSuccess) {  // core.hpp:383
    hwt_kernel<<<(a.batch + 63) / 64, 64, 0, stream>>>(a, hit, bitmap, used);
    hwt_repair_kernel<<<1, 256, 0, stream>>>(a, hit, bitmap, used, res);
    e = cudaGetLastError();
    if (e == cudaSuccess && result) e = cudaMemcpyAsync(result, res, sizeof(*result), cudaMemcpyDeviceToHost, stream);
  }
  cudaFreeAsync(scratch, stream);
  return e;
}

cudaError_t launch_sampler(int kind, const SampleArgs &a, int num_sms, cudaStream_t stream, cudaMemPool_t pool, unsigned long long *hwt_used) {
  const uint64_t total = (uint64_t)a.batch * a.blocks_per_poly;
  if (total == 0) return cudaSuccess;
  uint64_t blocks = (total + 255) / 256;
  if (blocks > (uint64_t)num_sms * 16) blocks = (uint64_t)num_sms * 16;
  switch (kind) {
    case SAMPLE_UNIFORM: uniform_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a); break;
    case SAMPLE_NON_UNIFORM: non_uniform_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a); break;
    case SAMPLE_ZO: zo_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a); break;
    case SAMPLE_HWT: return launch_hwt(a, stream, pool, hwt_used);
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace nflgpu

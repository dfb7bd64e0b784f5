// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// extern "C" driver around the UNMODIFIED reference Gaussian sampler (include/nfl/prng/FastGaussianNoise.hpp, compiled from
// /root/reference where it lies) and poly::set(gaussian) (core.hpp:284-336), built into oracle/_ref/libnflref.so.
// MPFR/GMP: the image has the runtimes (libmpfr.so.6, libgmp.so.10) without headers; oracle/shim/ declares the entry points.
//   * nflref_gaussian_create/info/barriers   the reference's own barrier table (cumulative distribution, MPFR arithmetic) —
//     read out of the private members, which is why this one translation unit sees FastGaussianNoise with `private`
//     spelled `public` (access specifiers do not change the layout; every other TU uses the header untouched);
//   * nflref_gaussian_sample                 successive poly::set(gaussian(&prng, amplifier)) draws with the fixed-key PRNG
//     of ref_harness.cpp, reporting the first nonce used and how many fastrandombytes calls (nonces) the draws consumed —
//     data dependent (getNoise refills its buffer, FastGaussianNoise.hpp:601-610), found by probing the keystream.
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <list>
#include <tuple>
#include <typeinfo>
#include <vector>
#include <gmp.h>
#include <mpfr.h>
#include "fastrandombytes.h"
#define private public
#include "nfl/prng/FastGaussianNoise.hpp"
#undef private
#include <nfl.hpp>
#include "nfl/prng/crypto_stream_salsa20.h"

extern "C" unsigned long long *nflref_nonce_counter(void);  // ref_harness.cpp part 0: mirrors fastrandombytes.cpp's nonce

namespace {

struct Handle {
  int in_bytes, depth, limb_bits;
  void *obj;
};
std::vector<Handle> g_handles;

template <class IN, class T, unsigned D> nfl::FastGaussianNoise<IN, T, D> *as(const Handle &h) {
  return static_cast<nfl::FastGaussianNoise<IN, T, D> *>(h.obj);
}

// the nonce fastrandombytes will use next: draw 16 bytes and find them in the keystreams of the fixed key
unsigned long long probe_nonce(unsigned long long from, unsigned long long span) {
  unsigned char got[16], key[32], cand[16], nonce[8];
  for (int i = 0; i < 32; ++i) key[i] = (unsigned char)(i + 1);
  nfl::fastrandombytes(got, 16);
  for (unsigned long long n = from; n <= from + span; ++n) {
    for (int i = 0; i < 8; ++i) nonce[i] = (unsigned char)(n >> (8 * i));
    nfl_crypto_stream_salsa20_amd64_xmm6(cand, 16, nonce, key);
    if (!memcmp(cand, got, 16)) return n;
  }
  return ~0ull;
}

template <class IN, class T, unsigned D, size_t N, size_t M>
int sample(const Handle &h, void *out, size_t batch, unsigned long long amplifier, unsigned long long *first_nonce,
           unsigned long long *nonces_used) {
  typedef nfl::poly<T, N, M> P;
  unsigned long long *ctr = nflref_nonce_counter();
  *first_nonce = *ctr;
  P *o = static_cast<P *>(out);
  for (size_t i = 0; i < batch; ++i) o[i].set(nfl::gaussian<IN, T, D>(as<IN, T, D>(h), amplifier));
  const unsigned long long next = probe_nonce(*ctr + batch, 8 * batch + 8);
  if (next == ~0ull) return -3;
  *nonces_used = next - *ctr;
  *ctr = next + 1;  // the probe itself consumed one
  return 0;
}

}  // namespace

extern "C" {

// in_bytes 1|2 (uint8_t / uint16_t look-up words), depth 1|2; limb_bits = sizeof(out_class)*8.  Returns a handle >= 0.
int nflref_gaussian_create(double sigma, unsigned security, unsigned samples, double center, int in_bytes, int depth, int limb_bits) {
  Handle h{in_bytes, depth, limb_bits, nullptr};
#define MK(IN, IB, T, LB, D) \
  if (in_bytes == IB && limb_bits == LB && depth == D) h.obj = new nfl::FastGaussianNoise<IN, T, D>(sigma, security, samples, center);
  MK(uint8_t, 1, uint64_t, 64, 2) MK(uint8_t, 1, uint64_t, 64, 1) MK(uint16_t, 2, uint64_t, 64, 1)
  MK(uint8_t, 1, uint32_t, 32, 2) MK(uint8_t, 1, uint32_t, 32, 1) MK(uint16_t, 2, uint32_t, 32, 1)
  MK(uint8_t, 1, uint16_t, 16, 2) MK(uint8_t, 1, uint16_t, 16, 1) MK(uint16_t, 2, uint16_t, 16, 1)
#undef MK
  if (!h.obj) return -1;
  g_handles.push_back(h);
  return (int)g_handles.size() - 1;
}

// the fields are identical for every out_class; read them through the <.., uint64_t, ..> view only when that is the real type
#define WITH(h, ...)                                                                                              \
  do {                                                                                                            \
    if (h.in_bytes == 1 && h.depth == 2 && h.limb_bits == 64) { auto *g = as<uint8_t, uint64_t, 2>(h); __VA_ARGS__; }    \
    else if (h.in_bytes == 1 && h.depth == 1 && h.limb_bits == 64) { auto *g = as<uint8_t, uint64_t, 1>(h); __VA_ARGS__; } \
    else if (h.in_bytes == 2 && h.depth == 1 && h.limb_bits == 64) { auto *g = as<uint16_t, uint64_t, 1>(h); __VA_ARGS__; } \
    else if (h.in_bytes == 1 && h.depth == 2 && h.limb_bits == 32) { auto *g = as<uint8_t, uint32_t, 2>(h); __VA_ARGS__; } \
    else if (h.in_bytes == 1 && h.depth == 1 && h.limb_bits == 32) { auto *g = as<uint8_t, uint32_t, 1>(h); __VA_ARGS__; } \
    else if (h.in_bytes == 2 && h.depth == 1 && h.limb_bits == 32) { auto *g = as<uint16_t, uint32_t, 1>(h); __VA_ARGS__; } \
    else if (h.in_bytes == 1 && h.depth == 2 && h.limb_bits == 16) { auto *g = as<uint8_t, uint16_t, 2>(h); __VA_ARGS__; } \
    else if (h.in_bytes == 1 && h.depth == 1 && h.limb_bits == 16) { auto *g = as<uint8_t, uint16_t, 1>(h); __VA_ARGS__; } \
    else if (h.in_bytes == 2 && h.depth == 1 && h.limb_bits == 16) { auto *g = as<uint16_t, uint16_t, 1>(h); __VA_ARGS__; } \
  } while (0)

// info[0..6] = number_of_barriers, word_precision, bit_precision, flag_ctr1, flag_ctr2, rounded_center (as int64), lu_size
int nflref_gaussian_info(int handle, long long *info, double *tail_bound) {
  if (handle < 0 || handle >= (int)g_handles.size()) return -1;
  const Handle &h = g_handles[handle];
  WITH(h, (info[0] = g->_number_of_barriers, info[1] = g->_word_precision, info[2] = g->_bit_precision, info[3] = g->_flag_ctr1,
           info[4] = g->_flag_ctr2, info[5] = g->rounded_center, info[6] = g->_lu_size, *tail_bound = g->_tail_bound));
  return 0;
}

// out: number_of_barriers * word_precision look-up words (in_bytes each), row-major, exactly as the reference holds them
int nflref_gaussian_barriers(int handle, void *out) {
  if (handle < 0 || handle >= (int)g_handles.size()) return -1;
  const Handle &h = g_handles[handle];
  WITH(h, {
    const size_t wp = g->_word_precision, row = wp * (size_t)h.in_bytes;
    for (unsigned i = 0; i < g->_number_of_barriers; ++i) memcpy(static_cast<char *>(out) + i * row, g->barriers[i], row);
  });
  return 0;
}

// out[0..batch) = successive poly::set(gaussian(&prng, amplifier)) draws (core.hpp:291-325)
int nflref_gaussian_sample(int handle, size_t degree, size_t nmoduli, void *out, size_t batch, unsigned long long amplifier,
                           unsigned long long *first_nonce, unsigned long long *nonces_used) {
  if (handle < 0 || handle >= (int)g_handles.size()) return -1;
  if (reinterpret_cast<uintptr_t>(out) & 31) return -2;
  const Handle &h = g_handles[handle];
#define CASE(IN, IB, D, T, LB, N, M) \
  if (h.in_bytes == IB && h.depth == D && h.limb_bits == LB && degree == N && nmoduli == M) \
    return sample<IN, T, D, N, M>(h, out, batch, amplifier, first_nonce, nonces_used);
  CASE(uint8_t, 1, 2, uint64_t, 64, 1024, 4) CASE(uint8_t, 1, 1, uint64_t, 64, 1024, 4) CASE(uint16_t, 2, 1, uint64_t, 64, 1024, 4)
  CASE(uint8_t, 1, 2, uint64_t, 64, 64, 3)
  CASE(uint8_t, 1, 2, uint32_t, 32, 4096, 1) CASE(uint16_t, 2, 1, uint32_t, 32, 4096, 1)
  CASE(uint8_t, 1, 2, uint16_t, 16, 512, 2) CASE(uint8_t, 1, 1, uint16_t, 16, 512, 2)
#undef CASE
  return -1;
}

}  // extern "C"

// Internal interface of the discrete-Gaussian sampler: host tables (gaussian.cpp) and device launcher (sampler.cu).
#ifndef NFLGPU_GAUSSIAN_H
#define NFLGPU_GAUSSIAN_H
#include <cstdint>
#include <vector>
#ifdef __CUDACC__
#include <cuda_runtime.h>
#endif

namespace nflgpu {

enum { GAUSS_MAX_ROW_BYTES = 64 };  // bytes of one barrier = precision of the cumulative distribution (512 bits)
enum { GAUSS_WALK_THREADS = 32, GAUSS_SMEM_BUDGET = 200 * 1024 };

// One look-up entry.  sub: -1 = not flagged; first level, depth 2: number of the second-level table (>= 1);
// otherwise 0 = flagged, walk barriers [bstart, bstart + bcount).
struct GaussLutEntry {
  int32_t val;
  int32_t sub;
  uint32_t bstart, bcount;
};

struct GaussianTable {
  unsigned nb = 0, wp = 0, bit_precision = 0, lu_size = 0, flag_ctr1 = 0, flag_ctr2 = 0;
  int in_bytes = 1, depth = 2;
  long rounded_center = 0;
  double tail_bound = 0;
  std::vector<unsigned char> barriers;  // [nb][wp] look-up words, most significant first, native byte order inside a word
  std::vector<GaussLutEntry> lut;       // table 0 = first level, tables 1.. = second level
};

bool gaussian_runtime_available();
int gaussian_compute_barriers(double sigma, unsigned security, unsigned samples, double center, int in_bytes, GaussianTable *t);
int gaussian_build_luts(GaussianTable *t, int depth);
uint64_t gaussian_words_per_fill(const GaussianTable &t, uint64_t degree);

#ifdef __CUDACC__  // the device side (capi.cu, sampler.cu); gaussian.cpp is plain host C++
struct GaussArgs {
  void *dst;
  const uint64_t *moduli;
  uint32_t key[8];
  uint64_t first_nonce, amplifier, poly_bytes;
  uint64_t words_per_fill;   // innoise_words of getNoise()
  uint32_t nmoduli, log2_degree, limb_bits, batch;
  uint32_t wp, in_bytes, depth, lu_size;
  const unsigned char *barriers;
  const GaussLutEntry *lut;
  // scratch.  rows = window + a few: every nonce first_nonce + r, r < rows, gets its keystream evaluated at every position
  uint32_t window, rows;
  int32_t *pos_val;      // [rows][words_per_fill] output of an evaluation starting at that position
  uint8_t *pos_adv;      // [rows][pitch] look-up words it consumes; pitch = words_per_fill rounded up to 16, base 16-byte aligned
  uint32_t *cand_idx;    // [degree][window] where in pos_val output k of the draw starting at nonce first_nonce + c was evaluated
  uint32_t *cand_calls;  // [window] fastrandombytes calls that draw makes
  uint32_t *chosen;      // [batch] candidate of polynomial b
  uint64_t *result;      // [0] = nonces consumed by the batch, [1] = 1 when the window was too small
  uint32_t walk_rows, walk_stride;  // set by the launcher
};
cudaError_t launch_gaussian(GaussArgs a, int device, int num_sms, cudaStream_t stream);
#endif

}  // namespace nflgpu
#endif

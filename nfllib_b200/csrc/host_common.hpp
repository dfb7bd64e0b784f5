// Host-side internals shared by the translation units of libnflgpu.so (not part of the public ABI).
#ifndef NFLGPU_HOST_COMMON_HPP
#define NFLGPU_HOST_COMMON_HPP

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace nflgpu {

struct LimbLimits {
  uint64_t kMaxPolyDegree;
  uint64_t kMaxNbModuli;
  unsigned kModulusBitsize;
};
bool limb_limits(int limb_bits, LimbLimits *out);
bool derive_params(int limb_bits, size_t first, size_t count, uint64_t *P, uint64_t *Pn, uint64_t *roots,
                   uint64_t *invkmax);
uint64_t newton_pn(int limb_bits, uint64_t p);
uint64_t powmod64(uint64_t b, uint64_t e, uint64_t m);
uint64_t invmod64(uint64_t a, uint64_t p);

// One residue's twiddle tables in the device layout described in ntt_plan.h, as {w, shoup_w} pairs widened
// to uint64_t (the uploader narrows them to the kernel's word type).
struct ResidueTables {
  std::vector<uint64_t> fwd_w, fwd_ws;  // N entries each
  std::vector<uint64_t> inv_w, inv_ws;  // plan_inv_entries() each (ntt_plan.h: N, or N + N/2 when N^-1 is folded into the twiddles); [N-1] = N^-1
};
// limb_bits: 16/32/64 (Shoup shift); word_bits: 32 or 64 (kernel word: 16-bit limbs compute in 32-bit words)
// raw = true builds the tables of the cyclic transform core::ntt / core::inv_ntt (no phi twist, no N^-1)
void build_residue_tables(int limb_bits, int word_bits, size_t N, uint64_t p, uint64_t root, uint64_t kmax,
                          ResidueTables *out, bool raw = false);

void set_error(const std::string &msg);

// pageable <-> pinned staging copy of the host-buffer pipeline (host_copy.cpp: copy-thread pool, non-temporal stores)
void staging_copy(void *dst, const void *src, size_t bytes);

}  // namespace nflgpu
#endif

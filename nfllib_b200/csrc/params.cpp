// NFLlib parameter tables, re-derived from their defining rules (host side, C++11).
//
// Replaces the data of include/nfl/params.hpp:12-119 + lib/params/params.cpp:1-17 without copying it:
//   P[i]      the primes 2^(w-2) - k*2*kMaxPolyDegree + 1, k = 1, 2, ..., in that (descending) order
//   Pn[i]     floor(2^(2w) / P[i]) - 2^(w+2)                      ("lower word of the Newton quotient")
//   roots[i]  g^((P[i]-1) / (2*kMaxPolyDegree)) with g the least primitive root of P[i]
//   invkmax   kMaxPolyDegree^-1 mod P[i]
// tests/test_params.py checks every entry (2 + 291 + 1000) against the reference's tables.
#include "host_common.hpp"

#include <algorithm>
#include <mutex>
#include <vector>

namespace nflgpu {

typedef unsigned __int128 u128;

static inline uint64_t mulmod64(uint64_t a, uint64_t b, uint64_t m) { return (uint64_t)(((u128)a * b) % m); }

uint64_t powmod64(uint64_t b, uint64_t e, uint64_t m) {
  uint64_t r = 1 % m;
  b %= m;
  while (e) {
    if (e & 1) r = mulmod64(r, b, m);
    b = mulmod64(b, b, m);
    e >>= 1;
  }
  return r;
}

// Deterministic Miller-Rabin for 64-bit integers (first twelve prime bases).
static bool is_prime64(uint64_t n) {
  if (n < 2) return false;
  static const uint64_t small[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  for (uint64_t q : small) {
    if (n == q) return true;
    if (n % q == 0) return false;
  }
  uint64_t d = n - 1;
  int s = 0;
  while (!(d & 1)) { d >>= 1; ++s; }
  for (uint64_t a : small) {
    uint64_t x = powmod64(a, d, n);
    if (x == 1 || x == n - 1) continue;
    bool composite = true;
    for (int i = 1; i < s; ++i) {
      x = mulmod64(x, x, n);
      if (x == n - 1) { composite = false; break; }
    }
    if (composite) return false;
  }
  return true;
}

static uint64_t gcd64(uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; }

// Pollard rho (Brent variant is unnecessary at this size); n composite, odd.
static uint64_t rho(uint64_t n) {
  for (uint64_t c = 1;; ++c) {
    uint64_t x = 2, y = 2, d = 1;
    while (d == 1) {
      x = (mulmod64(x, x, n) + c) % n;
      y = (mulmod64(y, y, n) + c) % n;
      y = (mulmod64(y, y, n) + c) % n;
      d = gcd64(x > y ? x - y : y - x, n);
    }
    if (d != n) return d;
  }
}

static void factor(uint64_t n, std::vector<uint64_t> &out) {
  if (n == 1) return;
  for (uint64_t q = 2; q < 64 && n > 1; ++q)
    if (n % q == 0) { out.push_back(q); while (n % q == 0) n /= q; }
  if (n == 1) return;
  if (is_prime64(n)) { out.push_back(n); return; }
  uint64_t d = rho(n);
  factor(d, out);
  factor(n / d, out);
}

static uint64_t least_primitive_root(uint64_t p) {
  std::vector<uint64_t> fs;
  factor(p - 1, fs);
  std::sort(fs.begin(), fs.end());
  fs.erase(std::unique(fs.begin(), fs.end()), fs.end());
  for (uint64_t g = 2;; ++g) {
    bool ok = true;
    for (uint64_t q : fs)
      if (powmod64(g, (p - 1) / q, p) == 1) { ok = false; break; }
    if (ok) return g;
  }
}

uint64_t invmod64(uint64_t a, uint64_t p) { return powmod64(a, p - 2, p); }

bool limb_limits(int limb_bits, LimbLimits *out) {
  // params.hpp:22-39 (uint16_t), 58-78 (uint32_t), 97-118 (uint64_t)
  switch (limb_bits) {
    case 16: *out = LimbLimits{512, 2, 14}; return true;
    case 32: *out = LimbLimits{32768, 291, 30}; return true;
    case 64: *out = LimbLimits{1048576, 1000, 62}; return true;
  }
  return false;
}

namespace {
struct Table {
  std::mutex mu;
  std::vector<uint64_t> P;      // primes found so far, in table order
  uint64_t next_k = 1;          // next multiplier to try
  std::vector<uint64_t> roots;  // lazily filled (0 = not yet computed)
};
Table g_tables[3];
int table_index(int bits) { return bits == 16 ? 0 : bits == 32 ? 1 : 2; }
}  // namespace

// Extends the prime list of `bits` up to index `upto` (exclusive) and returns it.
static bool ensure_primes(int bits, size_t upto) {
  LimbLimits lim;
  if (!limb_limits(bits, &lim) || upto > lim.kMaxNbModuli) return false;
  Table &t = g_tables[table_index(bits)];
  const uint64_t top = (uint64_t)1 << (bits - 2);
  const uint64_t step = 2 * lim.kMaxPolyDegree;
  while (t.P.size() < upto) {
    if (t.next_k * step >= top) return false;
    uint64_t cand = top - t.next_k * step + 1;
    ++t.next_k;
    if (is_prime64(cand)) { t.P.push_back(cand); t.roots.push_back(0); }
  }
  return true;
}

uint64_t newton_pn(int bits, uint64_t p) {
  // floor(2^(2w) / p) - 2^(w+2), kept to w bits.
  if (bits == 64) {
    u128 a = (u128)1 << 64;
    // 2^128 / p = q1 * 2^64 + q2 with q1 = floor(2^64 / p) >= 4 (p < 2^62); subtracting 2^66 = 4 * 2^64 and
    // keeping w = 64 bits leaves q2.
    uint64_t r1 = (uint64_t)(a % p);
    return (uint64_t)((((u128)r1) << 64) / p);
  }
  u128 q = ((u128)1 << (2 * bits)) / p - ((u128)1 << (bits + 2));
  return (uint64_t)q & (((uint64_t)1 << bits) - 1);
}

bool derive_params(int bits, size_t first, size_t count, uint64_t *P, uint64_t *Pn, uint64_t *roots,
                   uint64_t *invkmax) {
  LimbLimits lim;
  if (!limb_limits(bits, &lim)) return false;
  Table &t = g_tables[table_index(bits)];
  std::lock_guard<std::mutex> lock(t.mu);
  if (!ensure_primes(bits, first + count)) return false;
  for (size_t i = 0; i < count; ++i) {
    uint64_t p = t.P[first + i];
    if (P) P[i] = p;
    if (Pn) Pn[i] = newton_pn(bits, p);
    if (roots) {
      if (!t.roots[first + i]) {
        uint64_t g = least_primitive_root(p);
        t.roots[first + i] = powmod64(g, (p - 1) / (2 * lim.kMaxPolyDegree), p);
      }
      roots[i] = t.roots[first + i];
    }
    if (invkmax) invkmax[i] = invmod64(lim.kMaxPolyDegree % p, p);
  }
  return true;
}

}  // namespace nflgpu

// C ABI of libnflgpu.so (include/nflgpu.h): context, device buffers, launch wrappers, host-buffer pipeline.
// No CPU fallback anywhere: every compute entry point ends in a kernel launch or fails.
#include "../../include/nflgpu.h"
#include "host_common.hpp"
#include "gaussian.h"
#include "lift.h"
#include "ntt_dispatch.h"
#include "ntt_plan.h"
#include "pointwise.h"

#include <atomic>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace nflgpu {

static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }

cudaError_t launch_ntt(int limb_bits, int log2_degree, int mode, const NttLaunch &l, int device, int num_sms, cudaStream_t stream) {
  typedef cudaError_t (*fn_t)(int, const NttLaunch &, int, int, cudaStream_t);
  static const fn_t table[3][3] = {{launch_ntt_u16_fwd, launch_ntt_u16_inv, launch_ntt_u16_fwdmul},
                                   {launch_ntt_u32_fwd, launch_ntt_u32_inv, launch_ntt_u32_fwdmul},
                                   {launch_ntt_u64_fwd, launch_ntt_u64_inv, launch_ntt_u64_fwdmul}};
  const int li = limb_bits == 16 ? 0 : limb_bits == 32 ? 1 : limb_bits == 64 ? 2 : -1;
  if (li < 0 || mode < 0 || mode > 2) return cudaErrorInvalidValue;
  return table[li][mode](log2_degree, l, device, num_sms, stream);
}

bool ntt_supported(int limb_bits, int n) {
  switch (limb_bits) {
    case 64: return n >= 2 && n <= 20;
    case 32: return n >= 3 && n <= 15;
    case 16: return n >= 4 && n <= 9;
  }
  return false;
}

}  // namespace nflgpu

using namespace nflgpu;

#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t e_ = (expr);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      set_error(std::string(#expr) + ": " + cudaGetErrorName(e_) + " (" + cudaGetErrorString(e_) + ")"); \
      return NFLGPU_ERR_CUDA;                                                                            \
    }                                                                                                    \
  } while (0)

// Host-buffer pipeline (nflgpu_host_op / nflgpu_host_op_async): a ring of chunk slots fed by THREE streams -- every host->device
// copy goes out on `in`, every kernel on `run`, every device->host copy on `out`, chained per slot by events -- so that each
// copy engine sees one back-to-back queue (no idle gap while a per-chunk stream switches from its upload to its kernel to its
// download) and consecutive calls keep both PCIe directions busy across the call boundary.
struct HostSlot {
  void *dev[4] = {nullptr, nullptr, nullptr, nullptr};   // a, b, c, out
  void *pin[4] = {nullptr, nullptr, nullptr, nullptr};   // pinned staging (only used for pageable user buffers)
  cudaEvent_t up = nullptr, done = nullptr, down = nullptr;  // uploads finished / kernels finished / download finished
  bool busy = false;        // `down` has been recorded and not yet waited for
  void *unstage_to = nullptr;  // pageable destination of this slot's result (copied out of pin[3] when the slot retires)
  size_t unstage_bytes = 0;
};
struct HostPipe {
  static constexpr int kMaxRing = 16;
  cudaStream_t in = nullptr, run = nullptr, out = nullptr;
  HostSlot slot[kMaxRing];
  int ring = 0;            // slots in use (fixed at the first call: NFLGPU_HOST_RING, default 8)
  size_t slot_bytes = 0;   // capacity of every device / pinned buffer of a slot
  unsigned next = 0;       // next slot to use (slots retire in the order they were filled)
};

struct nflgpu_ctx {
  int limb_bits = 0, log2_degree = 0, device = 0, num_sms = 0;
  size_t degree = 0, nmoduli = 0, first_modulus = 0;
  size_t limb_bytes = 0;
  std::vector<uint64_t> moduli;
  void *d_moduli_word = nullptr;    // Word[nmoduli] for the NTT kernels
  uint64_t *d_moduli64 = nullptr;   // uint64_t[nmoduli] for the pointwise kernels
  uint64_t *d_consts = nullptr;     // Barrett constants, pointwise.h
  void *d_tw_fwd = nullptr, *d_tw_inv = nullptr;
  void *d_tw_raw_fwd = nullptr, *d_tw_raw_inv = nullptr;  // cyclic (no-twist) tables, built on first use
  std::vector<uint64_t> roots;
  uint64_t kmax = 0;
  std::mutex lazy_mu;  // guards the build-on-first-use tables below (raw twiddles, CRT lift)
  // CRT-lift tables, built on first use (lift.cu)
  int lift_words = 0;
  uint64_t *d_lift = nullptr;  // [inv | c64 | qhat (M*W) | q (W)]
  std::atomic<uint64_t> launches{0};
  // stream-ordered scratch (nflgpu_polymul temporaries, sampler scratch, nflgpu_scratch_alloc): a pool owned by the context,
  // destroyed with it, so that cached blocks never outlive the context or touch the application's default pool
  cudaMemPool_t pool = nullptr;
  // nflgpu_gather_residues: local slabs are placed on a side stream while the caller's stream pulls the peer slabs over NVLink
  cudaStream_t gather_stream = nullptr;
  cudaEvent_t gather_fork = nullptr, gather_join = nullptr;
  // dynamic unit scheduling of the NTT kernels (ntt_engine.cuh UnitWalk): one set of nmoduli + 1 device counters per
  // stream, created the first time the stream is seen.  Launches on one stream are ordered and every launch leaves its set
  // zeroed, so a set is never shared by two running kernels.  (A CUDA graph must be replayed on the stream it was captured
  // from, after one eager call on that stream has created the set.)
  std::mutex sched_mu;
  std::unordered_map<void *, uint32_t *> sched_by_stream;
  // The host-buffer pipeline uses per-context staging state: concurrent nflgpu_host_op calls on ONE context
  // must be serialised by the caller (distinct contexts are independent).
  HostPipe pipe;
};

namespace {

// builds the twiddle tables of all residues (tables.cpp), narrows them to the kernel word type and uploads them
int upload_tables(nflgpu_ctx *ctx, bool raw, void **d_fwd, void **d_inv) {
  const int word_bits = ctx->limb_bits == 64 ? 64 : 32;
  const size_t tw_entry = ctx->limb_bits == 64 ? 16 : 8, degree = ctx->degree, nmoduli = ctx->nmoduli;
  const size_t inv_entries = (size_t)plan_inv_entries((int)ctx->log2_degree, word_bits);  // N + N/2 when N^-1 is folded into the twiddles
  std::vector<unsigned char> hf(nmoduli * degree * tw_entry), hi(nmoduli * inv_entries * tw_entry);
  for (size_t cm = 0; cm < nmoduli; ++cm) {
    ResidueTables t;
    build_residue_tables(ctx->limb_bits, word_bits, degree, ctx->moduli[cm], ctx->roots[cm], ctx->kmax, &t, raw);
    for (size_t i = 0; i < degree; ++i) {
      if (word_bits == 64) {
        uint64_t *f = reinterpret_cast<uint64_t *>(hf.data()) + (cm * degree + i) * 2;
        f[0] = t.fwd_w[i]; f[1] = t.fwd_ws[i];
      } else {
        uint32_t *f = reinterpret_cast<uint32_t *>(hf.data()) + (cm * degree + i) * 2;
        f[0] = (uint32_t)t.fwd_w[i]; f[1] = (uint32_t)t.fwd_ws[i];
      }
    }
    for (size_t i = 0; i < inv_entries; ++i) {
      if (word_bits == 64) {
        uint64_t *v = reinterpret_cast<uint64_t *>(hi.data()) + (cm * inv_entries + i) * 2;
        v[0] = t.inv_w[i]; v[1] = t.inv_ws[i];
      } else {
        uint32_t *v = reinterpret_cast<uint32_t *>(hi.data()) + (cm * inv_entries + i) * 2;
        v[0] = (uint32_t)t.inv_w[i]; v[1] = (uint32_t)t.inv_ws[i];
      }
    }
  }
  cudaError_t e;
  if ((e = cudaMalloc(d_fwd, hf.size())) != cudaSuccess || (e = cudaMalloc(d_inv, hi.size())) != cudaSuccess ||
      (e = cudaMemcpy(*d_fwd, hf.data(), hf.size(), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(*d_inv, hi.data(), hi.size(), cudaMemcpyHostToDevice)) != cudaSuccess) {
    set_error(std::string("twiddle table upload: ") + cudaGetErrorName(e));
    return NFLGPU_ERR_CUDA;
  }
  return NFLGPU_OK;
}

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int check_buf(const nflgpu_ctx *ctx, const void *p, const char *name) {
  if (!ctx) { set_error("null context"); return NFLGPU_ERR_ARG; }
  if (!p) { set_error(std::string("null buffer: ") + name); return NFLGPU_ERR_ARG; }
  if (reinterpret_cast<uintptr_t>(p) & 15) { set_error(std::string("buffer not 16-byte aligned: ") + name); return NFLGPU_ERR_ARG; }
  return NFLGPU_OK;
}

int run_ntt(nflgpu_ctx *ctx, int mode, void *dst, const void *src, size_t batch, void *stream, const void *other = nullptr,
            bool raw = false) {
  int rc;
  if ((rc = check_buf(ctx, dst, "dst")) || (rc = check_buf(ctx, src, "src"))) return rc;
  if (mode == 2 && (rc = check_buf(ctx, other, "other"))) return rc;
  if (batch > 0xffffffffu) { set_error("batch too large"); return NFLGPU_ERR_ARG; }
  if (batch == 0) return NFLGPU_OK;
  DeviceGuard g(ctx->device);
  if (!g.ok) { set_error("cannot select CUDA device"); return NFLGPU_ERR_CUDA; }
  if (raw) {  // first use of the cyclic transform on this context builds its tables
    std::lock_guard<std::mutex> lock(ctx->lazy_mu);
    if (!ctx->d_tw_raw_fwd) {
      void *f = nullptr, *v = nullptr;
      if (int rc2 = upload_tables(ctx, true, &f, &v)) return rc2;
      ctx->d_tw_raw_inv = v;
      ctx->d_tw_raw_fwd = f;
    }
  }
  NttLaunch l;
  l.src = src; l.dst = dst; l.moduli = ctx->d_moduli_word;
  l.tw = raw ? (mode == 1 ? ctx->d_tw_raw_inv : ctx->d_tw_raw_fwd) : (mode == 1 ? ctx->d_tw_inv : ctx->d_tw_fwd);
  l.nmoduli = (uint32_t)ctx->nmoduli; l.batch = (uint32_t)batch;
  l.other = other; l.consts = ctx->d_consts;
  {
    std::lock_guard<std::mutex> lock(ctx->sched_mu);
    uint32_t *&set = ctx->sched_by_stream[stream];
    if (!set) {
      // zeroed on the launch stream itself: a cudaMemset on the legacy default stream is not ordered with a
      // non-blocking stream, and the first kernel would read the counters of a fresh allocation
      uint32_t *fresh = nullptr;
      CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&fresh), (ctx->nmoduli + 1) * sizeof(uint32_t)));
      cudaError_t e = cudaMemsetAsync(fresh, 0, (ctx->nmoduli + 1) * sizeof(uint32_t), (cudaStream_t)stream);
      if (e != cudaSuccess) {
        cudaFree(fresh);
        ctx->sched_by_stream.erase(stream);
        set_error(std::string("scheduler counters: ") + cudaGetErrorName(e));
        return NFLGPU_ERR_CUDA;
      }
      set = fresh;
    }
    l.sched = set;
  }
  CUDA_TRY(launch_ntt(ctx->limb_bits, ctx->log2_degree, mode, l, ctx->device, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  return NFLGPU_OK;
}

int run_pw(nflgpu_ctx *ctx, int op, int nin, void *dst, const void *a, const void *b, const void *c, const void *d, size_t batch,
           void *stream) {
  int rc;
  if ((rc = check_buf(ctx, dst, "dst")) || (rc = check_buf(ctx, a, "a"))) return rc;
  if (nin >= 2 && (rc = check_buf(ctx, b, "b"))) return rc;
  if (nin >= 3 && (rc = check_buf(ctx, c, "c"))) return rc;
  if (nin >= 4 && (rc = check_buf(ctx, d, "d"))) return rc;
  if (batch > 0xffffffffu) { set_error("batch too large"); return NFLGPU_ERR_ARG; }
  if (batch == 0) return NFLGPU_OK;
  DeviceGuard g(ctx->device);
  if (!g.ok) { set_error("cannot select CUDA device"); return NFLGPU_ERR_CUDA; }
  PwArgs p;
  p.dst = dst; p.a = a; p.b = b; p.c = c; p.d = d;
  p.moduli = ctx->d_moduli64; p.consts = ctx->d_consts;
  p.nmoduli = (uint32_t)ctx->nmoduli; p.degree = (uint32_t)ctx->degree; p.log2_degree = (uint32_t)ctx->log2_degree;
  p.batch = (uint32_t)batch;
  CUDA_TRY(launch_pointwise(ctx->limb_bits, op, p, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  return NFLGPU_OK;
}

}  // namespace

extern "C" {

const char *nflgpu_last_error(void) { return g_last_error.c_str(); }

int nflgpu_params_limits(int limb_bits, uint64_t *kmax, uint64_t *kmaxmod, unsigned *bits) {
  LimbLimits lim;
  if (!limb_limits(limb_bits, &lim)) { set_error("limb_bits must be 16, 32 or 64"); return NFLGPU_ERR_ARG; }
  if (kmax) *kmax = lim.kMaxPolyDegree;
  if (kmaxmod) *kmaxmod = lim.kMaxNbModuli;
  if (bits) *bits = lim.kModulusBitsize;
  return NFLGPU_OK;
}

int nflgpu_params(int limb_bits, size_t first, size_t count, uint64_t *P, uint64_t *Pn, uint64_t *roots, uint64_t *invkmax) {
  LimbLimits lim;
  if (!limb_limits(limb_bits, &lim)) { set_error("limb_bits must be 16, 32 or 64"); return NFLGPU_ERR_ARG; }
  if (first + count > lim.kMaxNbModuli) { set_error("modulus index beyond params<T>::kMaxNbModuli"); return NFLGPU_ERR_ARG; }
  if (!derive_params(limb_bits, first, count, P, Pn, roots, invkmax)) { set_error("parameter derivation failed"); return NFLGPU_ERR_ARG; }
  return NFLGPU_OK;
}

int nflgpu_ctx_create(nflgpu_ctx **out, int limb_bits, size_t degree, size_t nmoduli, size_t first_modulus, int device,
                      const uint64_t *moduli, const uint64_t *roots) {
  if (!out) { set_error("null ctx out-pointer"); return NFLGPU_ERR_ARG; }
  *out = nullptr;
  LimbLimits lim;
  if (!limb_limits(limb_bits, &lim)) { set_error("limb_bits must be 16, 32 or 64"); return NFLGPU_ERR_ARG; }
  int n = 0;
  while (((size_t)1 << n) < degree) ++n;
  if (degree == 0 || ((size_t)1 << n) != degree || degree > lim.kMaxPolyDegree) {
    set_error("degree must be a power of two <= params<T>::kMaxPolyDegree");  // core.hpp:55-60 static_asserts
    return NFLGPU_ERR_ARG;
  }
  if (nmoduli == 0 || ((moduli == nullptr) != (roots == nullptr))) { set_error("bad moduli/roots arguments"); return NFLGPU_ERR_ARG; }
  if (!moduli && first_modulus + nmoduli > lim.kMaxNbModuli) {
    set_error("nmoduli exceeds params<T>::kMaxNbModuli");  // core.hpp:57-58
    return NFLGPU_ERR_ARG;
  }
  if (!ntt_supported(limb_bits, n)) {
    set_error("unsupported (limb_bits, degree): kernels cover 2^2..2^20 (64-bit), 2^3..2^15 (32-bit), 2^4..2^9 (16-bit)");
    return NFLGPU_ERR_UNSUPPORTED;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("no CUDA device available (libnflgpu has no CPU fallback)");
    return NFLGPU_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { set_error("bad device ordinal"); return NFLGPU_ERR_ARG; }

  nflgpu_ctx *ctx = new (std::nothrow) nflgpu_ctx();
  if (!ctx) { set_error("out of host memory"); return NFLGPU_ERR_ALLOC; }
  ctx->limb_bits = limb_bits; ctx->degree = degree; ctx->log2_degree = n; ctx->nmoduli = nmoduli;
  ctx->first_modulus = first_modulus; ctx->device = device; ctx->limb_bytes = limb_bits / 8;
  ctx->moduli.resize(nmoduli);
  std::vector<uint64_t> rts(nmoduli);
  if (moduli) {
    for (size_t i = 0; i < nmoduli; ++i) { ctx->moduli[i] = moduli[i]; rts[i] = roots[i]; }
  } else if (!derive_params(limb_bits, first_modulus, nmoduli, ctx->moduli.data(), nullptr, rts.data(), nullptr)) {
    delete ctx; set_error("parameter derivation failed"); return NFLGPU_ERR_ARG;
  }
  const uint64_t beta4 = limb_bits == 64 ? ((uint64_t)1 << 62) : ((uint64_t)1 << (limb_bits - 2));
  for (size_t i = 0; i < nmoduli; ++i) {
    const uint64_t p = ctx->moduli[i];
    // lazy butterflies need 4p < 2^w; the Barrett constants assume floor(2^w / p) == 4 (params.hpp moduli are
    // just below 2^(w-2)); the root must have order exactly 2*kMax
    if (p >= beta4 || p <= beta4 / 5 * 4 || powmod64(rts[i], lim.kMaxPolyDegree, p) != p - 1) {
      delete ctx; set_error("modulus/root pair is not an NFLlib-style (w-2)-bit NTT prime"); return NFLGPU_ERR_ARG;
    }
  }

  DeviceGuard g(device);
  cudaDeviceProp prop;
  if (!g.ok || cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; set_error("cannot query CUDA device"); return NFLGPU_ERR_CUDA; }
  ctx->num_sms = prop.multiProcessorCount;

  // scratch comes from a pool of the context's own; freed blocks stay cached in it (with the default threshold 0 a
  // 768 MiB polymul scratch cost 15 ms per call) and are returned to the driver by nflgpu_ctx_trim / nflgpu_ctx_destroy
  {
    cudaMemPoolProps props;
    std::memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    if (cudaMemPoolCreate(&ctx->pool, &props) != cudaSuccess) {
      cudaGetLastError(); delete ctx; set_error("cannot create the context's memory pool"); return NFLGPU_ERR_CUDA;
    }
    uint64_t keep = ~0ull;
    cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  ctx->roots = rts;
  ctx->kmax = lim.kMaxPolyDegree;
  const int word_bits = limb_bits == 64 ? 64 : 32;
  std::vector<uint64_t> consts(nmoduli);
  std::vector<unsigned char> words(nmoduli * (word_bits / 8));
  for (size_t cm = 0; cm < nmoduli; ++cm) {
    const uint64_t p = ctx->moduli[cm];
    if (limb_bits == 64) { consts[cm] = newton_pn(64, p); reinterpret_cast<uint64_t *>(words.data())[cm] = p; }
    else {
      consts[cm] = limb_bits == 32 ? (uint64_t)((((unsigned __int128)1) << 64) / p) : 0;
      reinterpret_cast<uint32_t *>(words.data())[cm] = (uint32_t)p;
    }
  }
  if (int rc = upload_tables(ctx, false, &ctx->d_tw_fwd, &ctx->d_tw_inv)) { nflgpu_ctx_destroy(ctx); return rc; }
#define CTX_TRY(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      set_error(std::string(#expr) + ": " + cudaGetErrorName(e_));                            \
      nflgpu_ctx_destroy(ctx);                                                                \
      return NFLGPU_ERR_CUDA;                                                                 \
    }                                                                                         \
  } while (0)
  CTX_TRY(cudaMalloc(&ctx->d_moduli_word, words.size()));
  CTX_TRY(cudaMalloc(reinterpret_cast<void **>(&ctx->d_moduli64), nmoduli * 8));
  CTX_TRY(cudaMalloc(reinterpret_cast<void **>(&ctx->d_consts), nmoduli * 8));
  CTX_TRY(cudaMemcpy(ctx->d_moduli_word, words.data(), words.size(), cudaMemcpyHostToDevice));
  CTX_TRY(cudaMemcpy(ctx->d_moduli64, ctx->moduli.data(), nmoduli * 8, cudaMemcpyHostToDevice));
  CTX_TRY(cudaMemcpy(ctx->d_consts, consts.data(), nmoduli * 8, cudaMemcpyHostToDevice));
#undef CTX_TRY
  *out = ctx;
  return NFLGPU_OK;
}

int nflgpu_ctx_destroy(nflgpu_ctx *ctx) {
  if (!ctx) return NFLGPU_OK;
  DeviceGuard g(ctx->device);
  for (cudaStream_t st : {ctx->pipe.in, ctx->pipe.run, ctx->pipe.out}) if (st) cudaStreamSynchronize(st);
  for (auto &s : ctx->pipe.slot) {
    for (int i = 0; i < 4; ++i) { if (s.dev[i]) cudaFree(s.dev[i]); if (s.pin[i]) cudaFreeHost(s.pin[i]); }
    for (cudaEvent_t ev : {s.up, s.done, s.down}) if (ev) cudaEventDestroy(ev);
  }
  for (cudaStream_t st : {ctx->pipe.in, ctx->pipe.run, ctx->pipe.out}) if (st) cudaStreamDestroy(st);
  cudaFree(ctx->d_tw_fwd); cudaFree(ctx->d_tw_inv); cudaFree(ctx->d_tw_raw_fwd); cudaFree(ctx->d_tw_raw_inv); cudaFree(ctx->d_lift); cudaFree(ctx->d_moduli_word); cudaFree(ctx->d_moduli64); cudaFree(ctx->d_consts);
  for (auto &kv : ctx->sched_by_stream) cudaFree(kv.second);
  if (ctx->gather_stream) { cudaStreamSynchronize(ctx->gather_stream); cudaStreamDestroy(ctx->gather_stream); }
  if (ctx->gather_fork) cudaEventDestroy(ctx->gather_fork);
  if (ctx->gather_join) cudaEventDestroy(ctx->gather_join);
  if (ctx->pool) { cudaDeviceSynchronize(); cudaMemPoolDestroy(ctx->pool); }
  delete ctx;
  return NFLGPU_OK;
}

int nflgpu_ctx_info(const nflgpu_ctx *ctx, int *limb_bits, size_t *degree, size_t *nmoduli, int *device) {
  if (!ctx) { set_error("null context"); return NFLGPU_ERR_ARG; }
  if (limb_bits) *limb_bits = ctx->limb_bits;
  if (degree) *degree = ctx->degree;
  if (nmoduli) *nmoduli = ctx->nmoduli;
  if (device) *device = ctx->device;
  return NFLGPU_OK;
}

int nflgpu_ctx_moduli(const nflgpu_ctx *ctx, uint64_t *out) {
  if (!ctx || !out) { set_error("null argument"); return NFLGPU_ERR_ARG; }
  std::memcpy(out, ctx->moduli.data(), ctx->nmoduli * 8);
  return NFLGPU_OK;
}

uint64_t nflgpu_ctx_launch_count(const nflgpu_ctx *ctx) { return ctx ? ctx->launches.load() : 0; }

size_t nflgpu_batch_bytes(const nflgpu_ctx *ctx, size_t batch) { return ctx ? batch * ctx->nmoduli * ctx->degree * ctx->limb_bytes : 0; }

int nflgpu_alloc(nflgpu_ctx *ctx, size_t batch, void **dptr) {
  if (!ctx || !dptr) { set_error("null argument"); return NFLGPU_ERR_ARG; }
  DeviceGuard g(ctx->device);
  size_t bytes = nflgpu_batch_bytes(ctx, batch);
  if (cudaMalloc(dptr, bytes ? bytes : 16) != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc failed"); return NFLGPU_ERR_ALLOC; }
  return NFLGPU_OK;
}

int nflgpu_free(nflgpu_ctx *ctx, void *dptr) {
  if (!ctx) { set_error("null context"); return NFLGPU_ERR_ARG; }
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaFree(dptr));
  return NFLGPU_OK;
}

int nflgpu_scratch_alloc(nflgpu_ctx *ctx, size_t batch, void **dptr, void *stream) {
  if (!ctx || !dptr) { set_error("null argument"); return NFLGPU_ERR_ARG; }
  DeviceGuard g(ctx->device);
  const size_t bytes = nflgpu_batch_bytes(ctx, batch);
  if (cudaMallocFromPoolAsync(dptr, bytes ? bytes : 16, ctx->pool, (cudaStream_t)stream) != cudaSuccess) {
    cudaGetLastError(); set_error("cudaMallocFromPoolAsync failed"); return NFLGPU_ERR_ALLOC;
  }
  return NFLGPU_OK;
}

int nflgpu_scratch_free(nflgpu_ctx *ctx, void *dptr, void *stream) {
  if (!ctx) { set_error("null context"); return NFLGPU_ERR_ARG; }
  if (!dptr) return NFLGPU_OK;
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaFreeAsync(dptr, (cudaStream_t)stream));
  return NFLGPU_OK;
}

int nflgpu_ctx_trim(nflgpu_ctx *ctx) {
  if (!ctx) { set_error("null context"); return NFLGPU_ERR_ARG; }
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemPoolTrimTo(ctx->pool, 0));
  return NFLGPU_OK;
}

int nflgpu_upload(nflgpu_ctx *ctx, void *dst_dev, const void *src_host, size_t batch, void *stream) {
  if (!ctx || !dst_dev || !src_host) { set_error("null argument"); return NFLGPU_ERR_ARG; }
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaMemcpyAsync(dst_dev, src_host, nflgpu_batch_bytes(ctx, batch), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return NFLGPU_OK;
}

int nflgpu_download(nflgpu_ctx *ctx, void *dst_host, const void *src_dev, size_t batch, void *stream) {
  if (!ctx || !dst_host || !src_dev) { set_error("null argument"); return NFLGPU_ERR_ARG; }
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaMemcpyAsync(dst_host, src_dev, nflgpu_batch_bytes(ctx, batch), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return NFLGPU_OK;
}

int nflgpu_sync(nflgpu_ctx *ctx, void *stream) {
  if (!ctx) { set_error("null context"); return NFLGPU_ERR_ARG; }
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return NFLGPU_OK;
}

int nflgpu_ntt_fwd(nflgpu_ctx *ctx, void *dst, const void *src, size_t batch, void *stream) { return run_ntt(ctx, 0, dst, src, batch, stream); }
int nflgpu_ntt_inv(nflgpu_ctx *ctx, void *dst, const void *src, size_t batch, void *stream) { return run_ntt(ctx, 1, dst, src, batch, stream); }

int nflgpu_ntt_raw_fwd(nflgpu_ctx *ctx, void *dst, const void *src, size_t batch, void *stream) {
  return run_ntt(ctx, 0, dst, src, batch, stream, nullptr, true);
}
int nflgpu_ntt_raw_inv(nflgpu_ctx *ctx, void *dst, const void *src, size_t batch, void *stream) {
  return run_ntt(ctx, 1, dst, src, batch, stream, nullptr, true);
}

int nflgpu_mul(nflgpu_ctx *ctx, void *dst, const void *a, const void *b, size_t batch, void *stream) {
  return run_pw(ctx, PW_MUL, 2, dst, a, b, nullptr, nullptr, batch, stream);
}
int nflgpu_add(nflgpu_ctx *ctx, void *dst, const void *a, const void *b, size_t batch, void *stream) {
  return run_pw(ctx, PW_ADD, 2, dst, a, b, nullptr, nullptr, batch, stream);
}
int nflgpu_sub(nflgpu_ctx *ctx, void *dst, const void *a, const void *b, size_t batch, void *stream) {
  return run_pw(ctx, PW_SUB, 2, dst, a, b, nullptr, nullptr, batch, stream);
}
int nflgpu_mul_shoup(nflgpu_ctx *ctx, void *dst, const void *a, const void *b, const void *bprime, size_t batch, void *stream) {
  return run_pw(ctx, PW_MUL_SHOUP, 3, dst, a, b, bprime, nullptr, batch, stream);
}
int nflgpu_compute_shoup(nflgpu_ctx *ctx, void *dst, const void *a, size_t batch, void *stream) {
  return run_pw(ctx, PW_COMPUTE_SHOUP, 1, dst, a, nullptr, nullptr, nullptr, batch, stream);
}
int nflgpu_muladd(nflgpu_ctx *ctx, void *dst, const void *a, const void *b, const void *c, size_t batch, void *stream) {
  return run_pw(ctx, PW_MULADD, 3, dst, a, b, c, nullptr, batch, stream);
}
int nflgpu_muladd_shoup(nflgpu_ctx *ctx, void *dst, const void *a, const void *b, const void *c, const void *cprime, size_t batch,
                        void *stream) {
  return run_pw(ctx, PW_MULADD_SHOUP, 4, dst, a, b, c, cprime, batch, stream);
}

static int run_compare(nflgpu_ctx *ctx, bool want_equal, uint8_t *flags, const void *a, const void *b, size_t batch, void *stream) {
  int rc;
  if ((rc = check_buf(ctx, a, "a")) || (rc = check_buf(ctx, b, "b"))) return rc;
  if (!flags) { set_error("null flags buffer"); return NFLGPU_ERR_ARG; }
  if (batch > 0xffffffffu) { set_error("batch too large"); return NFLGPU_ERR_ARG; }
  if (batch == 0) return NFLGPU_OK;
  DeviceGuard g(ctx->device);
  if (!g.ok) { set_error("cannot select CUDA device"); return NFLGPU_ERR_CUDA; }
  CUDA_TRY(launch_compare(ctx->limb_bits, want_equal, a, b, flags, (uint32_t)batch, ctx->nmoduli * ctx->degree * ctx->limb_bytes, ctx->num_sms,
                          (cudaStream_t)stream));
  ctx->launches++;
  return NFLGPU_OK;
}
int nflgpu_any_eq(nflgpu_ctx *ctx, uint8_t *flags, const void *a, const void *b, size_t batch, void *stream) {
  return run_compare(ctx, true, flags, a, b, batch, stream);
}
int nflgpu_any_neq(nflgpu_ctx *ctx, uint8_t *flags, const void *a, const void *b, size_t batch, void *stream) {
  return run_compare(ctx, false, flags, a, b, batch, stream);
}

int nflgpu_eval(nflgpu_ctx *ctx, void *dst, const void *const *operands, size_t noperands, const uint8_t *program, size_t ntokens,
                size_t batch, void *stream) {
  int rc;
  if ((rc = check_buf(ctx, dst, "dst"))) return rc;
  if (!operands || !program || noperands == 0 || noperands > EV_MAX_OPERANDS || ntokens == 0 || ntokens > EV_MAX_TOKENS) {
    set_error("nflgpu_eval: need 1..8 operands and 1..32 tokens");
    return NFLGPU_ERR_ARG;
  }
  for (size_t i = 0; i < noperands; ++i)
    if ((rc = check_buf(ctx, operands[i], "operand"))) return rc;
  // validate the stack discipline on the host so the kernel never has to
  int sp = 0;
  for (size_t t = 0; t < ntokens; ++t) {
    const uint8_t tok = program[t];
    int pops, pushes = 1;
    if (tok < EV_MAX_OPERANDS) { if (tok >= noperands) { set_error("nflgpu_eval: operand index out of range"); return NFLGPU_ERR_ARG; } pops = 0; }
    else if (tok == EV_ADD || tok == EV_SUB || tok == EV_MUL) pops = 2;
    else if (tok == EV_MUL_SHOUP) pops = 3;
    else if (tok == EV_COMPUTE_SHOUP) pops = 1;
    else { set_error("nflgpu_eval: unknown token"); return NFLGPU_ERR_ARG; }
    if (sp < pops) { set_error("nflgpu_eval: stack underflow"); return NFLGPU_ERR_ARG; }
    sp += pushes - pops;
    if (sp > EV_MAX_STACK) { set_error("nflgpu_eval: expression too deep (stack > 8)"); return NFLGPU_ERR_ARG; }
  }
  if (sp != 1) { set_error("nflgpu_eval: program must leave exactly one value"); return NFLGPU_ERR_ARG; }
  if (batch > 0xffffffffu) { set_error("batch too large"); return NFLGPU_ERR_ARG; }
  if (batch == 0) return NFLGPU_OK;
  DeviceGuard g(ctx->device);
  if (!g.ok) { set_error("cannot select CUDA device"); return NFLGPU_ERR_CUDA; }
  EvArgs a;
  std::memset(&a, 0, sizeof(a));
  a.dst = dst;
  a.moduli = ctx->d_moduli64; a.consts = ctx->d_consts;
  a.nmoduli = (uint32_t)ctx->nmoduli; a.degree = (uint32_t)ctx->degree; a.log2_degree = (uint32_t)ctx->log2_degree;
  a.batch = (uint32_t)batch; a.ntokens = (uint32_t)ntokens;
  // Small trees have a kernel compiled for their program (eval_static.cu): canonical form = every leaf occurrence becomes a
  // leaf of its own, numbered in order of appearance (an operand used twice is simply passed twice)
  if (ntokens <= 7 && !std::getenv("NFLGPU_EVAL_INTERPRET")) {
    uint64_t key = 0;
    size_t nleaves = 0;
    for (size_t t = 0; t < ntokens; ++t) {
      uint8_t tok = program[t];
      if (tok < EV_MAX_OPERANDS) { a.operands[nleaves] = operands[tok]; tok = (uint8_t)nleaves++; }
      a.program[t] = tok;
      key |= (uint64_t)(tok + 1) << (8 * t);
    }
    cudaError_t e = cudaSuccess;
    if (launch_eval_static(ctx->limb_bits, key, a, ctx->num_sms, (cudaStream_t)stream, &e)) {
      CUDA_TRY(e);
      ctx->launches++;
      return NFLGPU_OK;
    }
    std::memset(a.operands, 0, sizeof(a.operands));
  }
  for (size_t i = 0; i < noperands; ++i) a.operands[i] = operands[i];
  std::memcpy(a.program, program, ntokens);
  CUDA_TRY(launch_eval(ctx->limb_bits, a, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  return NFLGPU_OK;
}

// ---- CRT lift (gmp.hpp:113-219) ----------------------------------------------------------------------------------
namespace {
typedef std::vector<uint64_t> Big;  // little-endian 64-bit words
Big big_mul_small(const Big &a, uint64_t m) {
  Big r(a.size() + 1, 0);
  unsigned __int128 carry = 0;
  for (size_t i = 0; i < a.size(); ++i) { carry += (unsigned __int128)a[i] * m; r[i] = (uint64_t)carry; carry >>= 64; }
  r[a.size()] = (uint64_t)carry;
  while (r.size() > 1 && r.back() == 0) r.pop_back();
  return r;
}
uint64_t big_mod_small(const Big &a, uint64_t m) {
  unsigned __int128 r = 0;
  for (size_t i = a.size(); i-- > 0;) r = ((r << 64) | a[i]) % m;
  return (uint64_t)r;
}
size_t big_bits(const Big &a) {
  size_t top = a.size() - 1;
  return a[top] ? top * 64 + (64 - __builtin_clzll(a[top])) : 0;
}
// moduli product, Q/p_cm, their inverses (gmp.hpp:116-150) -> device
int ensure_lift_tables(nflgpu_ctx *ctx) {
  std::lock_guard<std::mutex> lock(ctx->lazy_mu);
  if (ctx->d_lift) return NFLGPU_OK;
  const size_t M = ctx->nmoduli;
  Big Q(1, 1);
  for (uint64_t p : ctx->moduli) Q = big_mul_small(Q, p);
  const size_t W = (big_bits(Q) + 63) / 64;
  if (W > LIFT_MAX_WORDS) { set_error("CRT lift supports moduli products of at most 1024 bits"); return NFLGPU_ERR_UNSUPPORTED; }
  std::vector<uint64_t> host(2 * M + M * W + W, 0);
  for (size_t cm = 0; cm < M; ++cm) {
    Big qh(1, 1);
    for (size_t j = 0; j < M; ++j) if (j != cm) qh = big_mul_small(qh, ctx->moduli[j]);
    const uint64_t p = ctx->moduli[cm];
    host[cm] = invmod64(big_mod_small(qh, p), p);
    host[M + cm] = (uint64_t)((((unsigned __int128)1) << 64) % p);
    for (size_t k = 0; k < qh.size() && k < W; ++k) host[2 * M + cm * W + k] = qh[k];
  }
  for (size_t k = 0; k < Q.size() && k < W; ++k) host[2 * M + M * W + k] = Q[k];
  cudaError_t e;
  if ((e = cudaMalloc(reinterpret_cast<void **>(&ctx->d_lift), host.size() * 8)) != cudaSuccess ||
      (e = cudaMemcpy(ctx->d_lift, host.data(), host.size() * 8, cudaMemcpyHostToDevice)) != cudaSuccess) {
    set_error(std::string("lift table upload: ") + cudaGetErrorName(e));
    return NFLGPU_ERR_CUDA;
  }
  ctx->lift_words = (int)W;
  return NFLGPU_OK;
}
int run_lift(nflgpu_ctx *ctx, int dir, void *polys, uint64_t *words, size_t batch, void *stream, const void *const *slabs = nullptr,
             const size_t *first_residue = nullptr, const size_t *nresidues = nullptr, size_t nslabs = 0) {
  int rc;
  if (!slabs && (rc = check_buf(ctx, polys, "polys"))) return rc;
  if (ctx && ctx->nmoduli > LIFT_MAX_RESIDUES) { set_error("CRT lift supports at most 40 residues"); return NFLGPU_ERR_UNSUPPORTED; }
  if (!words || (reinterpret_cast<uintptr_t>(words) & 7)) { set_error("bad words buffer"); return NFLGPU_ERR_ARG; }
  if (batch > 0xffffffffu) { set_error("batch too large"); return NFLGPU_ERR_ARG; }
  if (batch == 0) return NFLGPU_OK;
  DeviceGuard g(ctx->device);
  if (!g.ok) { set_error("cannot select CUDA device"); return NFLGPU_ERR_CUDA; }
  if ((rc = ensure_lift_tables(ctx))) return rc;
  const size_t M = ctx->nmoduli, W = ctx->lift_words;
  LiftArgs a;
  a.polys = polys; a.words = words; a.moduli = ctx->d_moduli64; a.consts = ctx->d_consts;
  const size_t row = ctx->degree * ctx->limb_bytes;
  if (!slabs) {
    for (size_t cm = 0; cm < M; ++cm) { a.res_ptr[cm] = static_cast<const char *>(polys) + cm * row; a.res_stride[cm] = M * ctx->degree; }
  } else {  // residue groups in separate (possibly peer-device) slabs [batch][nres][degree]: every residue exactly once
    std::vector<int> seen(M, 0);
    for (size_t k = 0; k < nslabs; ++k) {
      if ((rc = check_buf(ctx, slabs[k], "slab"))) return rc;
      if (nresidues[k] == 0 || first_residue[k] + nresidues[k] > M) { set_error("lift: residue range outside the context"); return NFLGPU_ERR_ARG; }
      for (size_t j = 0; j < nresidues[k]; ++j) {
        const size_t cm = first_residue[k] + j;
        a.res_ptr[cm] = static_cast<const char *>(slabs[k]) + j * row;
        a.res_stride[cm] = nresidues[k] * ctx->degree;
        ++seen[cm];
      }
    }
    for (size_t cm = 0; cm < M; ++cm) if (seen[cm] != 1) { set_error("lift: the slabs must cover every residue exactly once"); return NFLGPU_ERR_ARG; }
  }
  a.inv = ctx->d_lift; a.c64 = ctx->d_lift + M; a.qhat = ctx->d_lift + 2 * M; a.q = ctx->d_lift + 2 * M + M * W;
  a.nmoduli = (uint32_t)M; a.log2_degree = (uint32_t)ctx->log2_degree; a.batch = (uint32_t)batch;
  CUDA_TRY(launch_lift(ctx->limb_bits, dir, (int)W, a, ctx->num_sms, (cudaStream_t)stream));
  ctx->launches++;
  return NFLGPU_OK;
}
}  // namespace

int nflgpu_lift_words(nflgpu_ctx *ctx, size_t *words_per_coefficient) {
  if (!ctx || !words_per_coefficient) { set_error("null argument"); return NFLGPU_ERR_ARG; }
  DeviceGuard g(ctx->device);
  if (int rc = ensure_lift_tables(ctx)) return rc;
  *words_per_coefficient = ctx->lift_words;
  return NFLGPU_OK;
}
int nflgpu_poly2mpz(nflgpu_ctx *ctx, uint64_t *dst_words, const void *src_polys, size_t batch, void *stream) {
  return run_lift(ctx, 0, const_cast<void *>(src_polys), dst_words, batch, stream);
}
int nflgpu_poly2mpz_slabs(nflgpu_ctx *ctx, uint64_t *dst_words, const void *const *slabs, const size_t *first_residue, const size_t *nresidues,
                          size_t nslabs, size_t batch, void *stream) {
  if (!ctx || !slabs || !first_residue || !nresidues || nslabs == 0) { set_error("nflgpu_poly2mpz_slabs: null argument"); return NFLGPU_ERR_ARG; }
  return run_lift(ctx, 0, nullptr, dst_words, batch, stream, slabs, first_residue, nresidues, nslabs);
}
int nflgpu_mpz2poly(nflgpu_ctx *ctx, void *dst_polys, const uint64_t *src_words, size_t batch, void *stream) {
  return run_lift(ctx, 1, dst_polys, const_cast<uint64_t *>(src_words), batch, stream);
}

static int run_sampler(nflgpu_ctx *ctx, int kind, void *dst, size_t batch, const uint8_t *key, uint64_t first_nonce, uint64_t p0,
                       uint64_t p1, uint64_t p2, size_t stream_bytes, void *stream, unsigned long long *hwt_used = nullptr) {
  int rc;
  if ((rc = check_buf(ctx, dst, "dst"))) return rc;
  if (!key) { set_error("null key"); return NFLGPU_ERR_ARG; }
  if (batch > 0xffffffffu) { set_error("batch too large"); return NFLGPU_ERR_ARG; }
  if (batch == 0) return NFLGPU_OK;
  DeviceGuard g(ctx->device);
  if (!g.ok) { set_error("cannot select CUDA device"); return NFLGPU_ERR_CUDA; }
  SampleArgs a;
  a.dst = dst; a.moduli = ctx->d_moduli64;
  for (int i = 0; i < 8; ++i)
    a.key[i] = (uint32_t)key[4 * i] | ((uint32_t)key[4 * i + 1] << 8) | ((uint32_t)key[4 * i + 2] << 16) | ((uint32_t)key[4 * i + 3] << 24);
  a.first_nonce = first_nonce;
  a.poly_bytes = ctx->nmoduli * ctx->degree * ctx->limb_bytes;
  a.blocks_per_poly = (stream_bytes + 63) / 64;  // keystream bytes one polynomial consumes
  a.nmoduli = (uint32_t)ctx->nmoduli; a.log2_degree = (uint32_t)ctx->log2_degree; a.limb_bits = (uint32_t)ctx->limb_bits;
  a.batch = (uint32_t)batch;
  a.param0 = p0; a.param1 = p1; a.param2 = p2;
  CUDA_TRY(launch_sampler(kind, a, ctx->num_sms, (cudaStream_t)stream, ctx->pool, hwt_used));
  if (kind == SAMPLE_HWT) ctx->launches++;  // the draw and the (normally idle) repair kernel
  ctx->launches++;
  return NFLGPU_OK;
}

int nflgpu_uniform(nflgpu_ctx *ctx, void *dst, size_t batch, const uint8_t key[32], uint64_t first_nonce, void *stream) {
  if (!ctx) { set_error("null context"); return NFLGPU_ERR_ARG; }
  return run_sampler(ctx, SAMPLE_UNIFORM, dst, batch, key, first_nonce, 0, 0, 0, ctx->nmoduli * ctx->degree * ctx->limb_bytes, stream);
}

int nflgpu_non_uniform(nflgpu_ctx *ctx, void *dst, size_t batch, uint64_t upper_bound, uint64_t amplifier, const uint8_t key[32],
                       uint64_t first_nonce, void *stream) {
  if (!ctx) { set_error("null context"); return NFLGPU_ERR_ARG; }
  if (upper_bound == 0) { set_error("upper_bound must be positive"); return NFLGPU_ERR_ARG; }
  for (uint64_t p : ctx->moduli)
    if (upper_bound >= p) { set_error("core: upper_bound is larger than the modulus"); return NFLGPU_ERR_ARG; }  // core.hpp:201-206
  // the reference computes the mask in double precision (core.hpp:218-219); reproduce that expression, not an integer log2
  const uint64_t mask = (1ULL << (int)(std::floor(std::log2((double)(2 * upper_bound - 1))) + 1)) - 1;
  return run_sampler(ctx, SAMPLE_NON_UNIFORM, dst, batch, key, first_nonce, upper_bound, amplifier, mask, ctx->degree * ctx->limb_bytes, stream);
}

int nflgpu_hwt_count(nflgpu_ctx *ctx, void *dst, size_t batch, uint32_t hwt, const uint8_t key[32], uint64_t first_nonce,
                     uint64_t *nonces_used, void *stream) {
  if (!ctx) { set_error("null context"); return NFLGPU_ERR_ARG; }
  if (hwt == 0 || hwt > ctx->degree) { set_error("hwt must be in [1, degree]"); return NFLGPU_ERR_ARG; }  // core.hpp:356
  const uint64_t calls = (ctx->degree - hwt + hwt - 1) / hwt + 1;  // refills of hwt words + the sign keystream
  uint64_t shift = 0;  // testing aid: shrink the rejection sampler's acceptance range so that tests can force extra refills
  if (const char *e = std::getenv("NFLGPU_HWT_TEST_REJECT_SHIFT")) shift = (uint64_t)std::strtoull(e, nullptr, 10) & 63;
  if (nonces_used) *nonces_used = 0;
  unsigned long long used = 0;
  int rc = run_sampler(ctx, SAMPLE_HWT, dst, batch, key, first_nonce, hwt, calls, shift, 64, stream, nonces_used ? &used : nullptr);
  if (rc != NFLGPU_OK || !nonces_used || batch == 0) return rc;
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  *nonces_used = used;
  return NFLGPU_OK;
}

int nflgpu_hwt(nflgpu_ctx *ctx, void *dst, size_t batch, uint32_t hwt, const uint8_t key[32], uint64_t first_nonce, void *stream) {
  return nflgpu_hwt_count(ctx, dst, batch, hwt, key, first_nonce, nullptr, stream);
}

int nflgpu_zo(nflgpu_ctx *ctx, void *dst, size_t batch, uint8_t rho, const uint8_t key[32], uint64_t first_nonce, void *stream) {
  if (!ctx) { set_error("null context"); return NFLGPU_ERR_ARG; }
  return run_sampler(ctx, SAMPLE_ZO, dst, batch, key, first_nonce, rho, 0, 0, ctx->degree, stream);
}

struct nflgpu_gaussian {
  nflgpu::GaussianTable t;
  int device = 0;
  unsigned char *d_barriers = nullptr;
  nflgpu::GaussLutEntry *d_lut = nullptr;
};

static int gaussian_finish(nflgpu_gaussian **out, nflgpu_ctx *ctx, nflgpu_gaussian *g, int lu_depth) {
  if (gaussian_build_luts(&g->t, lu_depth)) { delete g; return NFLGPU_ERR_UNSUPPORTED; }
  g->device = ctx->device;
  DeviceGuard dg(ctx->device);
  cudaError_t e = cudaErrorInvalidDevice;
  if (!dg.ok || (e = cudaMalloc(reinterpret_cast<void **>(&g->d_barriers), g->t.barriers.size())) != cudaSuccess ||
      (e = cudaMalloc(reinterpret_cast<void **>(&g->d_lut), g->t.lut.size() * sizeof(GaussLutEntry))) != cudaSuccess ||
      (e = cudaMemcpy(g->d_barriers, g->t.barriers.data(), g->t.barriers.size(), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(g->d_lut, g->t.lut.data(), g->t.lut.size() * sizeof(GaussLutEntry), cudaMemcpyHostToDevice)) != cudaSuccess) {
    set_error(std::string("Gaussian table upload: ") + cudaGetErrorName(e));
    nflgpu_gaussian_destroy(g);
    return NFLGPU_ERR_CUDA;
  }
  *out = g;
  return NFLGPU_OK;
}

int nflgpu_gaussian_create(nflgpu_gaussian **out, nflgpu_ctx *ctx, double sigma, unsigned security, unsigned samples, double center,
                           int in_bytes, int lu_depth) {
  if (!out || !ctx) { set_error("null argument"); return NFLGPU_ERR_ARG; }
  *out = nullptr;
  nflgpu_gaussian *g = new (std::nothrow) nflgpu_gaussian;
  if (!g) { set_error("out of memory"); return NFLGPU_ERR_ALLOC; }
  const int rc = gaussian_compute_barriers(sigma, security, samples, center, in_bytes, &g->t);
  if (rc) { delete g; return rc == -1 ? NFLGPU_ERR_ARG : NFLGPU_ERR_UNSUPPORTED; }
  return gaussian_finish(out, ctx, g, lu_depth);
}

int nflgpu_gaussian_create_from_barriers(nflgpu_gaussian **out, nflgpu_ctx *ctx, const void *barriers, size_t nbarriers,
                                         size_t word_precision, int in_bytes, int lu_depth, int64_t rounded_center) {
  if (!out || !ctx || !barriers) { set_error("null argument"); return NFLGPU_ERR_ARG; }
  *out = nullptr;
  if ((in_bytes != 1 && in_bytes != 2) || nbarriers == 0 || nbarriers > 0x7fffffffu || word_precision == 0 ||
      word_precision * in_bytes > GAUSS_MAX_ROW_BYTES) { set_error("bad Gaussian barrier table"); return NFLGPU_ERR_ARG; }
  nflgpu_gaussian *g = new (std::nothrow) nflgpu_gaussian;
  if (!g) { set_error("out of memory"); return NFLGPU_ERR_ALLOC; }
  g->t.nb = (unsigned)nbarriers; g->t.wp = (unsigned)word_precision; g->t.in_bytes = in_bytes;
  g->t.bit_precision = (unsigned)(word_precision * 8 * in_bytes); g->t.rounded_center = (long)rounded_center;
  g->t.barriers.assign(static_cast<const unsigned char *>(barriers), static_cast<const unsigned char *>(barriers) + nbarriers * word_precision * in_bytes);
  return gaussian_finish(out, ctx, g, lu_depth);
}

int nflgpu_gaussian_table(double sigma, unsigned security, unsigned samples, double center, int in_bytes, int lu_depth, int64_t info[7],
                          double *tail_bound, void *barriers, size_t capacity) {
  if (!info) { set_error("null argument"); return NFLGPU_ERR_ARG; }
  nflgpu_gaussian g;
  const int rc = gaussian_compute_barriers(sigma, security, samples, center, in_bytes, &g.t);
  if (rc) return rc == -1 ? NFLGPU_ERR_ARG : NFLGPU_ERR_UNSUPPORTED;
  if (gaussian_build_luts(&g.t, lu_depth)) return NFLGPU_ERR_UNSUPPORTED;
  nflgpu_gaussian_info(&g, info, tail_bound);
  if (barriers) {
    if (capacity < g.t.barriers.size()) { set_error("barrier buffer too small"); return NFLGPU_ERR_ARG; }
    std::memcpy(barriers, g.t.barriers.data(), g.t.barriers.size());
  }
  return NFLGPU_OK;
}

int nflgpu_gaussian_destroy(nflgpu_gaussian *g) {
  if (!g) return NFLGPU_OK;
  DeviceGuard dg(g->device);
  cudaFree(g->d_barriers); cudaFree(g->d_lut);
  delete g;
  return NFLGPU_OK;
}

int nflgpu_gaussian_info(const nflgpu_gaussian *g, int64_t info[7], double *tail_bound) {
  if (!g || !info) { set_error("null argument"); return NFLGPU_ERR_ARG; }
  info[0] = g->t.nb; info[1] = g->t.wp; info[2] = g->t.bit_precision; info[3] = g->t.flag_ctr1; info[4] = g->t.flag_ctr2;
  info[5] = g->t.rounded_center; info[6] = g->t.lu_size;
  if (tail_bound) *tail_bound = g->t.tail_bound;
  return NFLGPU_OK;
}

int nflgpu_gaussian_barriers(const nflgpu_gaussian *g, void *out) {
  if (!g || !out) { set_error("null argument"); return NFLGPU_ERR_ARG; }
  std::memcpy(out, g->t.barriers.data(), g->t.barriers.size());
  return NFLGPU_OK;
}

int nflgpu_gaussian_sample(nflgpu_ctx *ctx, const nflgpu_gaussian *g, void *dst, size_t batch, uint64_t amplifier, const uint8_t key[32],
                           uint64_t first_nonce, uint64_t *nonces_used, void *stream) {
  int rc;
  if ((rc = check_buf(ctx, dst, "dst"))) return rc;
  if (!g || !key) { set_error("null argument"); return NFLGPU_ERR_ARG; }
  if (g->device != ctx->device) { set_error("Gaussian sampler belongs to another device"); return NFLGPU_ERR_ARG; }
  if (batch > 0x3fffffffu) { set_error("batch too large"); return NFLGPU_ERR_ARG; }
  if (nonces_used) *nonces_used = 0;
  if (batch == 0) return NFLGPU_OK;
  const uint64_t words = gaussian_words_per_fill(g->t, ctx->degree);
  // getNoise() needs room for one full-precision comparison in a fresh buffer (the reference would read past its buffer)
  if (words <= 2ull * g->t.wp) { set_error("degree too small for this Gaussian sampler's refill logic"); return NFLGPU_ERR_UNSUPPORTED; }
  DeviceGuard dg(ctx->device);
  if (!dg.ok) { set_error("cannot select CUDA device"); return NFLGPU_ERR_CUDA; }
  cudaStream_t s = (cudaStream_t)stream;
  GaussArgs a;
  a.dst = dst; a.moduli = ctx->d_moduli64;
  for (int i = 0; i < 8; ++i)
    a.key[i] = (uint32_t)key[4 * i] | ((uint32_t)key[4 * i + 1] << 8) | ((uint32_t)key[4 * i + 2] << 16) | ((uint32_t)key[4 * i + 3] << 24);
  a.first_nonce = first_nonce; a.amplifier = amplifier;
  a.poly_bytes = ctx->nmoduli * ctx->degree * ctx->limb_bytes;
  a.words_per_fill = words;
  a.nmoduli = (uint32_t)ctx->nmoduli; a.log2_degree = (uint32_t)ctx->log2_degree; a.limb_bits = (uint32_t)ctx->limb_bits;
  a.batch = (uint32_t)batch;
  a.wp = g->t.wp; a.in_bytes = (uint32_t)g->t.in_bytes; a.depth = (uint32_t)g->t.depth; a.lu_size = g->t.lu_size;
  a.barriers = g->d_barriers; a.lut = g->d_lut;
  if (words * g->t.in_bytes + 64 > GAUSS_SMEM_BUDGET) { set_error("degree too large for the Gaussian sampler's shared-memory keystream"); return NFLGPU_ERR_UNSUPPORTED; }
  // Large batches go out in chunks that bound the scratch (the nonce chain simply continues from chunk to chunk).  Inside a
  // chunk a draw makes one call plus (usually at most) one refill: the window holds two nonces per polynomial (+16); if
  // the chain still runs out of it the chunk is halved.
  const size_t per_cand = ctx->degree * sizeof(int32_t) + 4 + (size_t)words * 5;
  size_t budget = (size_t)768 << 20;
  if (const char *mb = std::getenv("NFLGPU_GAUSS_SCRATCH_MB")) budget = (size_t)std::strtoull(mb, nullptr, 10) << 20;  // testing aid
  size_t chunk = budget / (2 * per_cand);
  if (chunk == 0) chunk = 1;
  if (chunk > 12000) chunk = 12000;  // the chain kernel keeps two jump tables of the window in shared memory
  uint64_t total_used = 0;
  for (size_t done = 0; done < batch;) {
    const size_t nb = batch - done < chunk ? batch - done : chunk;
    a.dst = static_cast<unsigned char *>(dst) + done * a.poly_bytes;
    a.batch = (uint32_t)nb;
    a.first_nonce = first_nonce + total_used;
    a.window = (uint32_t)(nb * 2 + 16);
    a.rows = a.window + 8;
    if ((uint64_t)a.rows * words >= (1ull << 32)) { chunk = nb / 2 ? nb / 2 : 1; if (nb == 1) break; continue; }  // 32-bit evaluation indices
    // scratch layout: result[2] | pos_val | cand_idx | cand_calls | chosen | pos_adv (rows of `pitch` bytes)
    const size_t val_bytes = ((size_t)a.rows * words * 4 + 15) & ~(size_t)15, idx_bytes = ((size_t)a.window * ctx->degree * 4 + 15) & ~(size_t)15;
    const size_t small_bytes = (((size_t)a.window * 4 + nb * 4) + 15) & ~(size_t)15, pitch = (words + 15) & ~(size_t)15;
    const size_t bytes = 16 + val_bytes + idx_bytes + small_bytes + (size_t)a.rows * pitch;
    unsigned char *scratch = nullptr;
    CUDA_TRY(cudaMallocFromPoolAsync(reinterpret_cast<void **>(&scratch), bytes, ctx->pool, s));
    a.result = reinterpret_cast<uint64_t *>(scratch);
    a.pos_val = reinterpret_cast<int32_t *>(scratch + 16);
    a.cand_idx = reinterpret_cast<uint32_t *>(scratch + 16 + val_bytes);
    a.cand_calls = reinterpret_cast<uint32_t *>(scratch + 16 + val_bytes + idx_bytes);
    a.chosen = a.cand_calls + a.window;
    a.pos_adv = scratch + 16 + val_bytes + idx_bytes + small_bytes;  // 16-byte aligned: every part before it is a multiple of 16
    cudaError_t e = launch_gaussian(a, ctx->device, ctx->num_sms, s);
    uint64_t result[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(result, a.result, sizeof(result), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFreeAsync(scratch, s);
    CUDA_TRY(e);
    ctx->launches += 4;
    if (result[1] != 0) {  // the chain left the window (more than two nonces per draw on average): retry with fewer polynomials
      if (nb == 1) break;
      chunk = nb / 2;
      continue;
    }
    total_used += result[0];
    done += nb;
    if (done == batch) {
      if (nonces_used) *nonces_used = total_used;
      return NFLGPU_OK;
    }
  }
  set_error("Gaussian sampler: refill chain did not fit the candidate window");
  return NFLGPU_ERR_UNSUPPORTED;
}

int nflgpu_polymul(nflgpu_ctx *ctx, void *dst, const void *a, const void *b, size_t batch, void *stream) {
  int rc;
  if ((rc = check_buf(ctx, dst, "dst")) || (rc = check_buf(ctx, a, "a")) || (rc = check_buf(ctx, b, "b"))) return rc;
  if (batch == 0) return NFLGPU_OK;
  DeviceGuard g(ctx->device);
  cudaStream_t s = (cudaStream_t)stream;
  void *tmp = nullptr;
  const size_t bytes = nflgpu_batch_bytes(ctx, batch);
  CUDA_TRY(cudaMallocFromPoolAsync(&tmp, bytes, ctx->pool, s));
  // three launches: tmp = ntt(a);  dst = ntt(b) * tmp (product fused into the forward kernel's copy-out);  dst = intt(dst)
  if ((rc = run_ntt(ctx, 0, tmp, a, batch, stream)) || (rc = run_ntt(ctx, 2, dst, b, batch, stream, tmp)) ||
      (rc = run_ntt(ctx, 1, dst, dst, batch, stream))) {
    cudaFreeAsync(tmp, s);
    return rc;
  }
  CUDA_TRY(cudaFreeAsync(tmp, s));
  return NFLGPU_OK;
}

// ---- residues sharded over devices: peer-memory gather -----------------------------------------

int nflgpu_ipc_export(nflgpu_ctx *ctx, const void *dptr, nflgpu_ipc_handle *out) {
  if (!ctx || !dptr || !out) { set_error("null argument"); return NFLGPU_ERR_ARG; }
  static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(out->bytes), "CUDA IPC handle does not fit nflgpu_ipc_handle");
  DeviceGuard g(ctx->device);
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, const_cast<void *>(dptr)));
  std::memset(out->bytes, 0, sizeof(out->bytes));
  std::memcpy(out->bytes, &h, sizeof(h));
  return NFLGPU_OK;
}

int nflgpu_ipc_open(nflgpu_ctx *ctx, const nflgpu_ipc_handle *handle, void **peer_ptr) {
  if (!ctx || !handle || !peer_ptr) { set_error("null argument"); return NFLGPU_ERR_ARG; }
  DeviceGuard g(ctx->device);
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle->bytes, sizeof(h));
  CUDA_TRY(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return NFLGPU_OK;
}

int nflgpu_ipc_close(nflgpu_ctx *ctx, void *peer_ptr) {
  if (!ctx) { set_error("null context"); return NFLGPU_ERR_ARG; }
  if (!peer_ptr) return NFLGPU_OK;
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaIpcCloseMemHandle(peer_ptr));
  return NFLGPU_OK;
}

int nflgpu_gather_residues(nflgpu_ctx *ctx, void *dst_full, const void *const *slabs, const size_t *first_residue, const size_t *nresidues,
                           size_t nslabs, size_t batch, void *stream) {
  int rc;
  if ((rc = check_buf(ctx, dst_full, "dst_full"))) return rc;
  if (!slabs || !first_residue || !nresidues || nslabs == 0) { set_error("nflgpu_gather_residues: null argument"); return NFLGPU_ERR_ARG; }
  const size_t row = ctx->degree * ctx->limb_bytes, full_pitch = ctx->nmoduli * row;
  for (size_t k = 0; k < nslabs; ++k) {
    if ((rc = check_buf(ctx, slabs[k], "slab"))) return rc;
    if (nresidues[k] == 0 || first_residue[k] + nresidues[k] > ctx->nmoduli) { set_error("nflgpu_gather_residues: residue range outside the context"); return NFLGPU_ERR_ARG; }
  }
  if (batch == 0) return NFLGPU_OK;
  DeviceGuard g(ctx->device);
  if (!g.ok) { set_error("cannot select CUDA device"); return NFLGPU_ERR_CUDA; }
  // one strided DMA per slab: `batch` rows of nres*N limbs, written where they belong in [batch][M][N] (no padded slabs, no
  // second interleave pass); for a slab in a peer device's memory the copy engine pulls it over NVLink.  Slabs of this
  // device go out on a side stream (forked from and joined back into the caller's stream), so the local placement overlaps
  // the peer transfers instead of queueing behind them.
  std::vector<char> local(nslabs, 0);
  size_t nlocal = 0;
  for (size_t k = 0; k < nslabs; ++k) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, slabs[k]) == cudaSuccess && at.type == cudaMemoryTypeDevice && at.device == ctx->device) { local[k] = 1; ++nlocal; }
    cudaGetLastError();
  }
  const bool fork = nlocal > 0 && nlocal < nslabs;
  if (fork) {
    if (!ctx->gather_stream) {
      CUDA_TRY(cudaStreamCreateWithFlags(&ctx->gather_stream, cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&ctx->gather_fork, cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&ctx->gather_join, cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventRecord(ctx->gather_fork, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamWaitEvent(ctx->gather_stream, ctx->gather_fork, 0));
  }
  for (size_t k = 0; k < nslabs; ++k) {
    const size_t width = nresidues[k] * row;
    cudaStream_t s = (fork && local[k]) ? ctx->gather_stream : (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpy2DAsync(static_cast<char *>(dst_full) + first_residue[k] * row, full_pitch, slabs[k], width, width, batch,
                               cudaMemcpyDeviceToDevice, s));
  }
  if (fork) {
    CUDA_TRY(cudaEventRecord(ctx->gather_join, ctx->gather_stream));
    CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, ctx->gather_join, 0));
  }
  return NFLGPU_OK;
}

// ---- host-buffer pipeline ------------------------------------------------------------------------------------

// (staging copies between pageable user memory and the pinned ring buffers: host_copy.cpp)

int nflgpu_host_register(nflgpu_ctx *ctx, void *host_ptr, size_t bytes) {
  if (!ctx || !host_ptr || bytes == 0) { set_error("null argument"); return NFLGPU_ERR_ARG; }
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaHostRegister(host_ptr, bytes, cudaHostRegisterDefault));
  return NFLGPU_OK;
}

int nflgpu_host_unregister(nflgpu_ctx *ctx, void *host_ptr) {
  if (!ctx || !host_ptr) { set_error("null argument"); return NFLGPU_ERR_ARG; }
  DeviceGuard g(ctx->device);
  CUDA_TRY(cudaHostUnregister(host_ptr));
  return NFLGPU_OK;
}

static bool is_pinned(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

namespace {

int host_dispatch(nflgpu_ctx *ctx, int op, void *const d[4], size_t cnt, void *st) {
  switch (op) {
    case 0: return nflgpu_ntt_fwd(ctx, d[3], d[0], cnt, st);
    case 1: return nflgpu_ntt_inv(ctx, d[3], d[0], cnt, st);
    case 2: return nflgpu_mul(ctx, d[3], d[0], d[1], cnt, st);
    case 3: return nflgpu_mul_shoup(ctx, d[3], d[0], d[1], d[2], cnt, st);
    case 4: return nflgpu_compute_shoup(ctx, d[3], d[0], cnt, st);
    case 5: return nflgpu_add(ctx, d[3], d[0], d[1], cnt, st);
    case 6: return nflgpu_sub(ctx, d[3], d[0], d[1], cnt, st);
    case 8: return nflgpu_polymul(ctx, d[3], d[0], d[1], cnt, st);
    case 9: return nflgpu_muladd(ctx, d[3], d[0], d[1], d[2], cnt, st);
    case 10: return nflgpu_ntt_raw_fwd(ctx, d[3], d[0], cnt, st);
    case 11: return nflgpu_ntt_raw_inv(ctx, d[3], d[0], cnt, st);
  }
  set_error("unknown op");
  return NFLGPU_ERR_ARG;
}

// waits for the slot's download; a result staged for a pageable destination is handed over now
int host_retire(HostSlot &s) {
  if (!s.busy) return NFLGPU_OK;
  s.busy = false;
  void *to = s.unstage_to;
  s.unstage_to = nullptr;
  CUDA_TRY(cudaEventSynchronize(s.down));
  if (to) staging_copy(to, s.pin[3], s.unstage_bytes);
  return NFLGPU_OK;
}

// retires every slot, oldest first
int host_drain(nflgpu_ctx *ctx) {
  HostPipe &p = ctx->pipe;
  int rc = NFLGPU_OK;
  for (int i = 0; i < p.ring; ++i) {
    const int r = host_retire(p.slot[(p.next + i) % p.ring]);
    if (r != NFLGPU_OK && rc == NFLGPU_OK) rc = r;
  }
  return rc;
}

// after a failure: nothing of the pipeline may still be touching the caller's buffers or the staging memory
void host_abort(nflgpu_ctx *ctx) {
  HostPipe &p = ctx->pipe;
  for (cudaStream_t st : {p.in, p.run, p.out}) if (st) cudaStreamSynchronize(st);
  for (auto &s : p.slot) { s.busy = false; s.unstage_to = nullptr; }
}

long env_long(const char *name, long lo, long hi, long dflt) {
  const char *e = std::getenv(name);
  if (!e) return dflt;
  const long v = std::atol(e);
  return v < lo || v > hi ? dflt : v;
}

int host_pipeline(nflgpu_ctx *ctx, int op, void *dst_host, const void *a_host, const void *b_host, const void *c_host, size_t batch,
                  bool wait) {
  if (!ctx || !dst_host || !a_host) { set_error("null argument"); return NFLGPU_ERR_ARG; }
  int nin;
  switch (op) {
    case 0: case 1: case 4: case 10: case 11: nin = 1; break;
    case 2: case 5: case 6: case 8: nin = 2; break;
    case 3: case 9: nin = 3; break;
    default: set_error("unknown op"); return NFLGPU_ERR_ARG;
  }
  const void *in[3] = {a_host, b_host, c_host};
  for (int i = 0; i < nin; ++i) if (!in[i]) { set_error("missing operand"); return NFLGPU_ERR_ARG; }
  DeviceGuard g(ctx->device);
  if (!g.ok) { set_error("cannot select CUDA device"); return NFLGPU_ERR_CUDA; }
  HostPipe &p = ctx->pipe;
  if (batch == 0) return wait ? host_drain(ctx) : NFLGPU_OK;
  if (!p.out) {  // (`out` is created last: a failure half way is retried by the next call)
    for (cudaStream_t *st : {&p.in, &p.run, &p.out})
      if (!*st) CUDA_TRY(cudaStreamCreateWithFlags(st, cudaStreamNonBlocking));
  }
  const size_t poly_bytes = ctx->nmoduli * ctx->degree * ctx->limb_bytes;
  bool pinned[4] = {is_pinned(a_host), nin >= 2 && is_pinned(b_host), nin >= 3 && is_pinned(c_host), is_pinned(dst_host)};

  // Optional zero-copy path (NFLGPU_HOST_ZEROCOPY=1): when every operand lives in pinned (device-mapped, UVA) host
  // memory the kernels read and write it directly over PCIe — no staging buffers.  Measured on B200 / PCIe gen5 it
  // moves 37 GB/s each way against 41 GB/s for the staged pipeline (profiles/r01e_variants.log), so it is off by default.
  {
    static const bool zc = env_long("NFLGPU_HOST_ZEROCOPY", 0, 1, 0) == 1;
    bool all_pinned = pinned[3];
    for (int i = 0; i < nin; ++i) all_pinned = all_pinned && pinned[i];
    if (all_pinned && zc) {
      void *dp[4] = {nullptr, nullptr, nullptr, nullptr};
      const void *hp[4] = {a_host, b_host, c_host, dst_host};
      for (int i = 0; i < 4; ++i)
        if (hp[i]) CUDA_TRY(cudaHostGetDevicePointer(&dp[i], const_cast<void *>(hp[i]), 0));
      const int rc = host_dispatch(ctx, op, dp, batch, p.run);
      if (rc != NFLGPU_OK) return rc;
      CUDA_TRY(cudaStreamSynchronize(p.run));
      return NFLGPU_OK;
    }
  }

  // Chunk size and ring depth, measured on B200 / PCIe gen5 (profiles/r02_variants.log, block 7): copies of 8 MiB and more run at
  // the rate of whole-array copies (48 GB/s each way at once), smaller ones lose 8-35 %, so short first and last chunks (a ramp
  // C/8, C/8, C/4, C/2, C ... was built and measured) do not pay; the ring only has to cover upload + kernel + download of one chunk.
  // Blocking calls use 16 MiB chunks (their first upload and last download overlap nothing, so they should be short); asynchronous
  // calls stream, pay no fill or drain, and use 32 MiB chunks: the copy engines slow down while a transform kernel runs, and fewer,
  // fuller launches spend less time in kernels (streamed step: 5.92 ms with 16 MiB chunks, 5.72 ms with 32 MiB, 5.53 ms without kernels).
  static const size_t chunk_bytes = (size_t)env_long("NFLGPU_HOST_CHUNK_MIB", 1, 1024, 16) << 20;
  static const size_t chunk_bytes_async = (size_t)env_long("NFLGPU_HOST_ASYNC_CHUNK_MIB", 1, 1024, env_long("NFLGPU_HOST_CHUNK_MIB", 1, 1024, 32)) << 20;
  static const int ring_want = (int)env_long("NFLGPU_HOST_RING", 2, HostPipe::kMaxRing, 4);
  size_t chunk = (wait ? chunk_bytes : chunk_bytes_async) / poly_bytes;
  if (chunk == 0) chunk = 1;
  if (chunk > batch) chunk = batch;
  if (p.slot_bytes < chunk * poly_bytes || p.ring == 0) {  // (re)build the ring; buffers only ever grow
    int rc = host_drain(ctx);
    if (rc != NFLGPU_OK) { host_abort(ctx); return rc; }
    for (cudaStream_t st : {p.in, p.run, p.out}) CUDA_TRY(cudaStreamSynchronize(st));
    p.slot_bytes = 0;  // a failure below leaves "no ring": the next call builds all of it again
    p.ring = 0;
    for (int k = 0; k < ring_want; ++k) {
      HostSlot &s = p.slot[k];
      for (int i = 0; i < 4; ++i) {
        if (s.dev[i]) { void *old = s.dev[i]; s.dev[i] = nullptr; CUDA_TRY(cudaFree(old)); }
        if (s.pin[i]) { void *old = s.pin[i]; s.pin[i] = nullptr; CUDA_TRY(cudaFreeHost(old)); }
        CUDA_TRY(cudaMalloc(&s.dev[i], chunk * poly_bytes));
      }
      for (cudaEvent_t *ev : {&s.up, &s.done, &s.down})
        if (!*ev) CUDA_TRY(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
    }
    p.slot_bytes = chunk * poly_bytes;
    p.ring = ring_want;
    p.next = 0;
  }

#define PIPE_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t e_ = (expr);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      set_error(std::string(#expr) + ": " + cudaGetErrorName(e_) + " (" + cudaGetErrorString(e_) + ")"); \
      host_abort(ctx);                                                                                   \
      return NFLGPU_ERR_CUDA;                                                                            \
    }                                                                                                    \
  } while (0)
  // a call that fits one chunk has nothing to overlap: all three steps go on one stream (no event hops on the latency path)
  const bool single = batch <= chunk;
  // ... and a SMALL one (a single host poly, what poly::ntt_pow_phi() is) skips the copy engines altogether: the kernel reads its
  // operands from, and writes its result to, mapped pinned memory over PCIe -- one launch and one wait instead of upload, launch,
  // download and wait.  Measured (profiles/r02_variants.log blocks 12, 13): faster than the copy engines for every one-chunk size tried,
  // 32 KiB (26 vs 33 us) to 8 MiB (306 vs 353+ us) -- a lone chunk cannot overlap its upload with its download, the kernel's own
  // reads and writes do.  NFLGPU_HOST_SMALL_KIB sets the limit per operand (default 8 MiB, 0 = never).
  static const size_t small_bytes = (size_t)env_long("NFLGPU_HOST_SMALL_KIB", 0, 65536, 8192) << 10;
  // (not for the split transforms, N > 2^15: their global-memory passes work in place on the destination several times)
  bool direct = single && batch * poly_bytes <= small_bytes && ctx->log2_degree <= 15;
  for (int i = 0; i < nin; ++i) direct = direct && (!pinned[i] || (reinterpret_cast<uintptr_t>(in[i]) & 15) == 0);  // (kernels want 16-byte alignment;
  direct = direct && (!pinned[3] || (reinterpret_cast<uintptr_t>(dst_host) & 15) == 0);                              //  the copy engines do not care)
  cudaStream_t sin = single ? p.run : p.in, sout = single ? p.run : p.out;
  for (size_t done = 0; done < batch;) {
    HostSlot &s = p.slot[p.next];
    int rc = host_retire(s);  // the chunk that used this slot `ring` chunks ago
    if (rc != NFLGPU_OK) { host_abort(ctx); return rc; }
    const size_t cnt = (batch - done < chunk) ? batch - done : chunk;
    const size_t bytes = cnt * poly_bytes;
    void *kd[4] = {s.dev[0], s.dev[1], s.dev[2], s.dev[3]};  // what the kernels read and write
    for (int i = 0; i < nin; ++i) {
      const char *src = static_cast<const char *>(in[i]) + done * poly_bytes;
      if (!pinned[i]) {
        if (!s.pin[i]) PIPE_TRY(cudaHostAlloc(&s.pin[i], p.slot_bytes, cudaHostAllocDefault));
        staging_copy(s.pin[i], src, bytes);
        src = static_cast<const char *>(s.pin[i]);
      }
      if (direct) PIPE_TRY(cudaHostGetDevicePointer(&kd[i], const_cast<char *>(src), 0));
      else PIPE_TRY(cudaMemcpyAsync(s.dev[i], src, bytes, cudaMemcpyHostToDevice, sin));
    }
    char *user = static_cast<char *>(dst_host) + done * poly_bytes, *out = user;
    if (!pinned[3]) {
      if (!s.pin[3]) PIPE_TRY(cudaHostAlloc(&s.pin[3], p.slot_bytes, cudaHostAllocDefault));
      out = static_cast<char *>(s.pin[3]);
    }
    if (direct) PIPE_TRY(cudaHostGetDevicePointer(&kd[3], out, 0));
    if (!single) {
      PIPE_TRY(cudaEventRecord(s.up, sin));
      PIPE_TRY(cudaStreamWaitEvent(p.run, s.up, 0));
    }
#ifdef NFLGPU_HOST_NOKERNEL  // timing experiment (tools/e2e_sweep.py): copies and events only, results are wrong
    rc = NFLGPU_OK;
#else
    rc = host_dispatch(ctx, op, kd, cnt, p.run);
#endif
    if (rc != NFLGPU_OK) { host_abort(ctx); return rc; }
    if (!single) {
      PIPE_TRY(cudaEventRecord(s.done, p.run));
      PIPE_TRY(cudaStreamWaitEvent(sout, s.done, 0));
    }
    if (!direct) PIPE_TRY(cudaMemcpyAsync(out, s.dev[3], bytes, cudaMemcpyDeviceToHost, sout));
    PIPE_TRY(cudaEventRecord(s.down, sout));
    s.busy = true;
    s.unstage_to = pinned[3] ? nullptr : user;
    s.unstage_bytes = bytes;
    p.next = (p.next + 1) % (unsigned)p.ring;
    done += cnt;
  }
#undef PIPE_TRY
  if (wait) {
    const int rc = host_drain(ctx);
    if (rc != NFLGPU_OK) host_abort(ctx);
    return rc;
  }
  return NFLGPU_OK;
}

}  // namespace

int nflgpu_host_op(nflgpu_ctx *ctx, int op, void *dst_host, const void *a_host, const void *b_host, const void *c_host, size_t batch) {
  return host_pipeline(ctx, op, dst_host, a_host, b_host, c_host, batch, true);
}

int nflgpu_host_op_async(nflgpu_ctx *ctx, int op, void *dst_host, const void *a_host, const void *b_host, const void *c_host,
                         size_t batch) {
  return host_pipeline(ctx, op, dst_host, a_host, b_host, c_host, batch, false);
}

int nflgpu_host_sync(nflgpu_ctx *ctx) {
  if (!ctx) { set_error("null context"); return NFLGPU_ERR_ARG; }
  DeviceGuard g(ctx->device);
  if (!g.ok) { set_error("cannot select CUDA device"); return NFLGPU_ERR_CUDA; }
  const int rc = host_drain(ctx);
  if (rc != NFLGPU_OK) host_abort(ctx);
  return rc;
}

}  // extern "C"

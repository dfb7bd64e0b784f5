// Staging copies of the host-buffer pipeline (nflgpu_host_op on pageable memory): pageable user array <-> pinned ring buffer.
//
// One host thread moves ~8-10 GB/s, far below the PCIe rate the ring is fed at, and the pipeline's pageable path is bound by
// exactly these copies (two per chunk).  So: a small pool of persistent copy threads (no thread creation per chunk), each
// moving a disjoint slice, and non-temporal stores where the CPU has AVX2 -- the destination of a staging copy is either read
// next by the DMA engine or far larger than the caches, so write-allocating it only doubles the store traffic.
// NFLGPU_HOST_COPY_THREADS sets the pool size (default min(12, three quarters of the hardware threads); 1 = copy on the calling thread).
#include "host_common.hpp"

#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace nflgpu {

namespace {

#if defined(__x86_64__)
__attribute__((target("avx2"))) void copy_stream_avx2(char *dst, const char *src, size_t bytes) {
  // head: up to the first 32-byte boundary of dst
  size_t head = (32 - ((uintptr_t)dst & 31)) & 31;
  if (head > bytes) head = bytes;
  std::memcpy(dst, src, head);
  dst += head; src += head; bytes -= head;
  size_t blocks = bytes / 128;
  for (size_t i = 0; i < blocks; ++i) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + 64));
    const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst), a);
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + 96), d);
    src += 128; dst += 128;
  }
  _mm_sfence();
  std::memcpy(dst, src, bytes - blocks * 128);
}
bool have_avx2() {
  static const bool v = __builtin_cpu_supports("avx2");
  return v;
}
#endif

void copy_slice(char *dst, const char *src, size_t bytes) {
#if defined(__x86_64__)
  if (bytes >= 4096 && have_avx2()) { copy_stream_avx2(dst, src, bytes); return; }
#endif
  std::memcpy(dst, src, bytes);
}

class CopyPool {
 public:
  explicit CopyPool(unsigned workers) : nworkers_(workers) {
    for (unsigned i = 0; i < workers; ++i) threads_[i] = std::thread([this, i] { run(i); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> l(mu_);
      stop_ = true;
    }
    wake_.notify_all();
    for (unsigned i = 0; i < nworkers_; ++i) threads_[i].join();
  }
  unsigned workers() const { return nworkers_; }
  // splits [0, bytes) into workers + 1 page-aligned slices; the caller copies the first one
  void copy(char *dst, const char *src, size_t bytes) {
    const unsigned parts = nworkers_ + 1;
    const size_t slice = ((bytes / parts) + 4095) & ~(size_t)4095;
    {
      std::lock_guard<std::mutex> l(mu_);
      dst_ = dst; src_ = src; bytes_ = bytes; slice_ = slice;
      pending_ = nworkers_;
      ++generation_;
    }
    wake_.notify_all();
    copy_slice(dst, src, slice < bytes ? slice : bytes);
    std::unique_lock<std::mutex> l(mu_);
    done_.wait(l, [this] { return pending_ == 0; });
  }

 private:
  void run(unsigned index) {
    unsigned long long seen = 0;
    for (;;) {
      char *dst; const char *src; size_t bytes, slice;
      {
        std::unique_lock<std::mutex> l(mu_);
        wake_.wait(l, [&] { return stop_ || generation_ != seen; });
        if (stop_) return;
        seen = generation_;
        dst = dst_; src = src_; bytes = bytes_; slice = slice_;
      }
      const size_t off = (size_t)(index + 1) * slice;
      if (off < bytes) copy_slice(dst + off, src + off, bytes - off < slice ? bytes - off : slice);
      {
        std::lock_guard<std::mutex> l(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  static constexpr unsigned kMax = 15;
  std::thread threads_[kMax];
  unsigned nworkers_;
  std::mutex mu_;
  std::condition_variable wake_, done_;
  char *dst_ = nullptr;
  const char *src_ = nullptr;
  size_t bytes_ = 0, slice_ = 0;
  unsigned pending_ = 0;
  unsigned long long generation_ = 0;
  bool stop_ = false;
};

}  // namespace

void staging_copy(void *dst, const void *src, size_t bytes) {
  static const unsigned want = [] {
    const char *e = std::getenv("NFLGPU_HOST_COPY_THREADS");
    // three quarters of the hardware threads, at most 12 (measured on the 16-thread B200 host, profiles/r02_variants.log block 16:
    // 4 / 8 / 12 / 16 threads = 0.66 / 0.85-0.89 / 1.00 / 0.96 M transforms/s for blocking calls on pageable C2 batches)
    long hw = (long)std::thread::hardware_concurrency() * 3 / 4;
    long v = e ? std::atol(e) : (hw < 2 ? 2 : hw > 12 ? 12 : hw);
    return (unsigned)(v < 1 ? 1 : v > 16 ? 16 : v);
  }();
  if (bytes < ((size_t)1 << 20) || want == 1) {
    copy_slice(static_cast<char *>(dst), static_cast<const char *>(src), bytes);
    return;
  }
  // one pool per process, created on first use; calls are serialised (the pipeline copies one chunk at a time per context,
  // and two contexts staging at once would only fight for the same memory bandwidth)
  static std::mutex pool_mu;
  static CopyPool *pool = nullptr;
  std::lock_guard<std::mutex> l(pool_mu);
  if (!pool) pool = new CopyPool(want - 1);  // (lives until process exit: its threads sleep on a condition variable)
  pool->copy(static_cast<char *>(dst), static_cast<const char *>(src), bytes);
}

}  // namespace nflgpu

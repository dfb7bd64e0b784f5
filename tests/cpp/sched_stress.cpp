// Stress of the NTT kernels' unit scheduler without torch, so that compute-sanitizer --tool racecheck / memcheck can run it
// (racecheck aborts inside torch's CUDA start-up on this image): hundreds of back-to-back forward / inverse launches of the
// dynamically scheduled sizes, alternating between two non-blocking streams, small and ragged batches.  Checks through the C ABI
// only: inverse(forward(x)) == x for every launch pair and forward(x) identical on both streams.  Exit status 0 = all equal.
// build: see tests/cpp/Makefile (g++ + libcudart for the streams); run: compute-sanitizer --tool racecheck tests/cpp/sched_stress
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "nflgpu.h"

#define CHECK(x) do { int rc_ = (x); if (rc_ != 0) { std::fprintf(stderr, "%s -> %d (%s)\n", #x, rc_, nflgpu_last_error()); return 1; } } while (0)

static int run(int bits, size_t N, size_t M, size_t batch, int launches) {
  if (launches < 2) launches = 2;  // both streams must have produced a result before they are compared
  nflgpu_ctx *ctx = nullptr;
  CHECK(nflgpu_ctx_create(&ctx, bits, N, M, 0, 0, nullptr, nullptr));
  std::vector<uint64_t> P(M);
  CHECK(nflgpu_ctx_moduli(ctx, P.data()));
  const size_t limb = bits / 8, bytes = nflgpu_batch_bytes(ctx, batch);
  std::vector<unsigned char> host(bytes), got(bytes), fwd0(bytes);
  uint64_t s = 0x9E3779B97F4A7C15ull + N + batch;
  for (size_t b = 0; b < batch; ++b)
    for (size_t cm = 0; cm < M; ++cm)
      for (size_t i = 0; i < N; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const uint64_t v = s % P[cm];
        std::memcpy(&host[((b * M + cm) * N + i) * limb], &v, limb);
      }
  cudaStream_t st[2];
  if (cudaStreamCreateWithFlags(&st[0], cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&st[1], cudaStreamNonBlocking) != cudaSuccess) return 1;
  void *src, *f[2], *r[2];
  CHECK(nflgpu_alloc(ctx, batch, &src));
  for (int k = 0; k < 2; ++k) { CHECK(nflgpu_alloc(ctx, batch, &f[k])); CHECK(nflgpu_alloc(ctx, batch, &r[k])); }
  CHECK(nflgpu_upload(ctx, src, host.data(), batch, nullptr));
  CHECK(nflgpu_sync(ctx, nullptr));
  for (int i = 0; i < launches; ++i) {
    const int k = i & 1;
    CHECK(nflgpu_ntt_fwd(ctx, f[k], src, batch, st[k]));
    CHECK(nflgpu_ntt_inv(ctx, r[k], f[k], batch, st[k]));
  }
  int bad = 0;
  for (int k = 0; k < 2; ++k) {
    CHECK(nflgpu_sync(ctx, st[k]));
    CHECK(nflgpu_download(ctx, got.data(), r[k], batch, nullptr));
    CHECK(nflgpu_sync(ctx, nullptr));
    if (std::memcmp(got.data(), host.data(), bytes) != 0) { std::fprintf(stderr, "round trip differs (stream %d)\n", k); bad = 1; }
    CHECK(nflgpu_download(ctx, k == 0 ? fwd0.data() : got.data(), f[k], batch, nullptr));
    CHECK(nflgpu_sync(ctx, nullptr));
  }
  if (std::memcmp(got.data(), fwd0.data(), bytes) != 0) { std::fprintf(stderr, "forward results of the two streams differ\n"); bad = 1; }
  std::printf("u%d N=%zu M=%zu batch=%zu: %d x (fwd, inv) on two streams: %s\n", bits, N, M, batch, launches, bad ? "MISMATCH" : "ok");
  for (int k = 0; k < 2; ++k) { nflgpu_free(ctx, f[k]); nflgpu_free(ctx, r[k]); cudaStreamDestroy(st[k]); }
  nflgpu_free(ctx, src);
  nflgpu_ctx_destroy(ctx);
  return bad;
}

// The host-buffer ring (nflgpu_host_op_async / nflgpu_host_sync: three streams, events, pinned staging) on pageable arrays: two
// independent batches in flight, forward then inverse, more chunks than ring slots.
static int run_host(int bits, size_t N, size_t M, size_t batch) {
  nflgpu_ctx *ctx = nullptr;
  CHECK(nflgpu_ctx_create(&ctx, bits, N, M, 0, 0, nullptr, nullptr));
  std::vector<uint64_t> P(M);
  CHECK(nflgpu_ctx_moduli(ctx, P.data()));
  const size_t limb = bits / 8, bytes = nflgpu_batch_bytes(ctx, batch);
  std::vector<unsigned char> a(bytes), c(bytes), fa(bytes), fc(bytes), ra(bytes), rc(bytes);
  uint64_t s = 0x9E3779B97F4A7C15ull + N + batch;
  for (std::vector<unsigned char> *h : {&a, &c})
    for (size_t b = 0; b < batch; ++b)
      for (size_t cm = 0; cm < M; ++cm)
        for (size_t i = 0; i < N; ++i) {
          s ^= s << 13; s ^= s >> 7; s ^= s << 17;
          const uint64_t v = s % P[cm];
          std::memcpy(&(*h)[((b * M + cm) * N + i) * limb], &v, limb);
        }
  CHECK(nflgpu_host_op_async(ctx, 0, fa.data(), a.data(), nullptr, nullptr, batch));
  CHECK(nflgpu_host_op_async(ctx, 0, fc.data(), c.data(), nullptr, nullptr, batch));
  CHECK(nflgpu_host_sync(ctx));
  CHECK(nflgpu_host_op_async(ctx, 1, ra.data(), fa.data(), nullptr, nullptr, batch));
  CHECK(nflgpu_host_op_async(ctx, 1, rc.data(), fc.data(), nullptr, nullptr, batch));
  CHECK(nflgpu_host_op(ctx, 1, fc.data(), fc.data(), nullptr, nullptr, batch));  // blocking, in place, behind the asynchronous ones
  int bad = std::memcmp(ra.data(), a.data(), bytes) != 0 || std::memcmp(rc.data(), c.data(), bytes) != 0 || std::memcmp(fc.data(), c.data(), bytes) != 0;
  // one polynomial: the small-call path (the kernel reads and writes mapped pinned memory), forward then inverse in place
  const size_t one = nflgpu_batch_bytes(ctx, 1);
  std::vector<unsigned char> x(a.begin() + one, a.begin() + 2 * one);
  CHECK(nflgpu_host_op(ctx, 0, x.data(), x.data(), nullptr, nullptr, 1));
  bad |= std::memcmp(x.data(), fa.data() + one, one) != 0;
  CHECK(nflgpu_host_op(ctx, 1, x.data(), x.data(), nullptr, nullptr, 1));
  bad |= std::memcmp(x.data(), a.data() + one, one) != 0;
  std::printf("u%d N=%zu M=%zu batch=%zu: host ring, 2 x (fwd, inv) asynchronous + 1 blocking in place: %s\n", bits, N, M, batch, bad ? "MISMATCH" : "ok");
  nflgpu_ctx_destroy(ctx);
  return bad;
}

int main(int argc, char **argv) {
  const int launches = argc > 1 ? std::atoi(argv[1]) : 300;
  int bad = 0;
  if (argc > 2 && std::atoi(argv[2]) == 32768) {  // only the cluster kernels, with several units per cluster (pipelined inverse)
    bad |= run(64, 32768, 2, 200, launches / 4 + 1);
    return bad;
  }
  // more units than unit slots: the software-pipelined kernels (ntt_engine.cuh PIPE_INV / PIPE_FWD) prefetch their next unit
  bad |= run(32, 4096, 2, 1500, launches / 8 + 1);
  bad |= run(64, 4096, 2, 500, launches / 8 + 1);
  bad |= run(64, 8192, 1, 700, launches / 8 + 1);
  bad |= run(64, 16384, 1, 300, launches / 8 + 1);  // 128-bit pass-0 window (NttCfg::ADJ) + pipelined inverse
  bad |= run_host(64, 1024, 4, 2600);
  bad |= run(64, 2048, 3, 37, launches);     // dynamic unit walk, two-pass tile
  bad |= run(64, 2048, 3, 1, launches);
  bad |= run(32, 4096, 2, 3, launches);
  bad |= run(64, 32768, 1, 3, launches / 4);  // split transform: global-memory pass + tile kernel
  bad |= run(64, 1024, 4, 37, launches);      // static walk, named barriers
  bad |= run(16, 512, 2, 5, launches);
  return bad;
}

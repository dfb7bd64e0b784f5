// Issue ceiling of the butterfly code itself (round 2): the kernels' own fwd_pass / inv_pass (ntt_engine.cuh) applied over and
// over to a register window, twiddles from shared memory, NO global memory, NO exchange, NO barriers.  What it prints is the
// cost of the compiled instruction mix when nothing but instruction issue can limit it: cycles per warp-butterfly per
// sub-partition, to set beside the full kernels' figure (kernel time * 1.965 GHz * 592 sub-partitions / warp-butterflies).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../nfllib_b200/csrc -o bfly_ceiling bfly_ceiling.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "ntt_engine.cuh"
using namespace nflgpu;

template <int LB, int LOGN, int PASS, bool INV, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) ceiling_kernel(const typename NttCfg<LB, LOGN>::TW *twg, typename NttCfg<LB, LOGN>::Word p,
                                                                typename NttCfg<LB, LOGN>::Word *out, int iters) {
  typedef NttCfg<LB, LOGN> C;
  typedef typename C::Word Word;
  typedef typename C::TW TW;
  extern __shared__ __align__(16) unsigned char smem[];
  TW *tws = reinterpret_cast<TW *>(smem);
  for (int i = threadIdx.x; i < (INV ? C::INV_TW : C::N); i += blockDim.x) tws[i] = twg[i];  // (inverse: the N^-1-scaled half too, NttCfg::FOLD)
  __syncthreads();
  const Word twop = 2 * p, np = opaque_neg(p);
  const TW ninv = tws[C::N - 1];
  Word x[C::E];
#pragma unroll
  for (int k = 0; k < C::E; ++k) x[k] = (Word)(threadIdx.x * 2654435761u + k * 40503u + blockIdx.x) % p;
  const int tid0 = threadIdx.x % C::TPU;
  for (int it = 0; it < iters; ++it) {
    const int tid = (tid0 + it) & (C::TPU - 1);  // the twiddle addresses change, so the loads stay inside the loop
    if (INV) inv_pass<C, PASS>(x, pass_tw<C, PASS>(tws, tid), p, np, twop, ninv);
    else fwd_pass<C, PASS>(x, pass_tw<C, PASS>(tws, tid), np, twop);
  }
  Word s = 0;
#pragma unroll
  for (int k = 0; k < C::E; ++k) s += x[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int LB, int LOGN, int PASS, bool INV, int THREADS, int MINB> void run(const char *name) {
  typedef NttCfg<LB, LOGN> C;
  typedef typename C::Word Word;
  typedef typename C::TW TW;
  const Word p = LB == 64 ? (Word)4611686018326724609ull : (Word)1073479681u;
  std::vector<TW> h(C::INV_TW);
  uint64_t s = 88172645463325252ull;
  for (int i = 0; i < C::INV_TW; ++i) {  // any (w, floor(w * 2^w / p)) pairs will do for timing
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    const Word w = (Word)(s % p);
    h[i].x = w;
    h[i].y = (Word)((((unsigned __int128)w) << C::WB) / p);
  }
  TW *tw; Word *out;
  cudaMalloc(&tw, sizeof(TW) * C::INV_TW); cudaMalloc(&out, sizeof(Word) * 148 * MINB * THREADS);
  cudaMemcpy(tw, h.data(), sizeof(TW) * C::INV_TW, cudaMemcpyHostToDevice);
  auto k = ceiling_kernel<LB, LOGN, PASS, INV, THREADS, MINB>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(TW) * C::INV_TW));
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, THREADS, sizeof(TW) * C::INV_TW);
  const int iters = 2000;
  k<<<148 * MINB, THREADS, sizeof(TW) * C::INV_TW>>>(tw, p, out, 10);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<<<148 * MINB, THREADS, sizeof(TW) * C::INV_TW>>>(tw, p, out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  constexpr int r = plan_r(C::n, C::WB, PASS);
  const double bf_per_thread = (double)iters * r * (C::E / 2);
  const double warps_per_smsp = (double)MINB * THREADS / 32 / 4;
  const double cyc = ms * 1e-3 * 1.965e9 / (bf_per_thread * warps_per_smsp);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
  printf("%-58s %3d regs, %d CTAs x %4d threads/SM (occupancy query %d): %7.2f cycles per warp-butterfly per sub-partition  [%s]\n", name, fa.numRegs, MINB,
         THREADS, occ, cyc, cudaGetErrorString(cudaGetLastError()));
  cudaFree(tw); cudaFree(out);
}

int main() {
  run<64, 10, 1, false, 512, 2>("u64 N=1024 forward pass 1 (4 stages, top-bit range)");
  run<64, 10, 1, false, 1024, 1>("u64 N=1024 forward pass 1, one CTA of 1024");
  run<64, 10, 1, false, 256, 2>("u64 N=1024 forward pass 1, 4 warps per sub-partition");
  run<64, 10, 1, true, 512, 2>("u64 N=1024 inverse pass 1 (4 stages)");
  run<64, 13, 1, false, 256, 2>("u64 N=8192 forward pass 1 (5 stages, 32 coefficients)");
  run<64, 13, 1, true, 256, 2>("u64 N=8192 inverse pass 1");
  run<32, 12, 1, false, 256, 4>("u32 N=4096 forward pass 1 (4 stages)");
  run<32, 12, 1, true, 256, 4>("u32 N=4096 inverse pass 1");
  return 0;
}

// Integer-pipe micro-benchmarks for sm_100a (development aid): issue rates of the instructions the 64-bit Shoup
// butterfly is made of, alone and mixed, to calibrate the integer-pipe roofline quoted in DESIGN.md.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench ubench.cu ; run: ./ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

template <int MODE> __global__ void __launch_bounds__(256, 2) k(u64 *out, u32 a, u32 b, int iters) {
  u64 acc[8]; u32 lo[8], hi[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i] = threadIdx.x + i; lo[i] = threadIdx.x * 3 + i; hi[i] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
// every op feeds its own result back as a multiplicand / operand, so nothing can be hoisted or strength-reduced
#define WIDE(i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((u32)acc[i]), "r"(b));
#define MADLO(i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo[i]) : "r"(b), "r"(a));
#define MADHI(i) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(lo[i]) : "r"(b), "r"(a));
#define ADD64(i) asm volatile("add.cc.u32 %0, %0, %1; addc.u32 %1, %1, %0;" : "+r"(lo[i]), "+r"(hi[i]));
#define ADD32(i) asm volatile("add.u32 %0, %0, %1;" : "+r"(hi[i]) : "r"(lo[i]));
#define LOP(i) asm volatile("xor.b32 %0, %0, %1;" : "+r"(hi[i]) : "r"(lo[i]));
      if (MODE == 0) { REP8(WIDE) }
      if (MODE == 1) { REP8(MADLO) }
      if (MODE == 2) { REP8(MADHI) }
      if (MODE == 3) { REP8(ADD64) }
      if (MODE == 4) { REP8(ADD32) }
      if (MODE == 5) { REP8(WIDE) REP8(ADD32) }            // 1 WIDE : 1 ALU
      if (MODE == 6) { REP8(MADLO) REP8(ADD32) }           // 1 IMAD : 1 ALU
      if (MODE == 7) { REP8(WIDE) REP8(ADD32) REP8(LOP) }  // 1 WIDE : 2 ALU
      if (MODE == 8) { REP8(WIDE) REP8(MADLO) REP8(ADD32) REP8(LOP) REP8(ADD64) }  // 1 WIDE + 1 IMAD : 4 ALU
      if (MODE == 9) { REP8(MADLO) REP8(MADHI) }
    }
  }
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i] + lo[i] + hi[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char *name, double instr_per_iter, u64 *out) {
  int iters = 20000;
  int grid = 148 * 2;
  k<MODE><<<grid, 256>>>(out, 3, 5, 100);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(out, 3, 5, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double warp_instr = (double)grid * 8 * iters * instr_per_iter;   // per warp: 4 unrolls x 8 x ops, folded in instr_per_iter
  double per_sm_per_clk = warp_instr / 148.0 / (ms * 1e-3 * clk_khz * 1e3);
  printf("%-44s %8.3f ms  %6.3f warp-instr/clk/SM  (%5.3f per SMSP) at nominal %d MHz\n", name, ms, per_sm_per_clk, per_sm_per_clk / 4, clk_khz / 1000);
}

int main() {
  u64 *out; cudaMalloc(&out, 148 * 2 * 256 * 8);
  run<0>("IMAD.WIDE.U32 (64-bit acc)", 32, out);
  run<1>("IMAD lo", 32, out);
  run<2>("IMAD.HI", 32, out);
  run<3>("IADD3 + IADD3.X (64-bit add)", 64, out);
  run<4>("IADD3 (32-bit add)", 32, out);
  run<5>("1 WIDE : 1 IADD3", 64, out);
  run<6>("1 IMAD : 1 IADD3", 64, out);
  run<7>("1 WIDE : 1 IADD3 : 1 LOP3", 96, out);
  run<8>("WIDE + IMAD + IADD3 + LOP3 + ADD64(2)", 192, out);
  run<9>("IMAD lo + IMAD.HI", 64, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}

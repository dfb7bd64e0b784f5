// Integer-pipe issue-rate micro-benchmarks for sm_100a: how fast can one SM retire the instructions a 64-bit Shoup
// butterfly is made of?  Each thread runs ILP independent dependency chains whose multiplicands depend on the previous
// result (so ptxas can neither hoist nor strength-reduce them); 8 warps per sub-partition hide the latencies.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o intpipe intpipe.cu ; run: ./intpipe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;

constexpr int ILP = 8;

template <int MODE> __global__ void __launch_bounds__(256, 4) k(u64 *out, const u32 *seed, int iters) {
  u64 acc[ILP]; u32 a[ILP], b[ILP], c[ILP], d[ILP]; float f[ILP], g[ILP];
  const u32 predsrc = seed[threadIdx.x & 255] & 1;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { acc[i] = seed[(threadIdx.x + i) & 255]; a[i] = seed[(threadIdx.x * 7 + i) & 255] | 1; b[i] = seed[(threadIdx.x * 3 + i) & 255] | 3; c[i] = seed[(threadIdx.x * 5 + i) & 255]; d[i] = seed[(threadIdx.x * 11 + i) & 255]; f[i] = (float)(a[i] & 1023) * 1e-3f; g[i] = 1.0f + (float)(b[i] & 7) * 1e-6f; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (MODE == 0) {  // IMAD.WIDE.U32 with 64-bit accumulate, multiplicand = low word of the accumulator
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((u32)acc[i]), "r"(b[i]));
        } else if (MODE == 1) {  // IMAD (low 32 bits)
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(a[(i + 1) % ILP]));
        } else if (MODE == 2) {  // IMAD.HI
          asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(a[(i + 1) % ILP]));
        } else if (MODE == 3) {  // IADD3 (three-input add, ALU pipe)
          asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a[i]) : "r"(b[i]), "r"(a[(i + 1) % ILP]));
        } else if (MODE == 4) {  // 1 WIDE : 2 IADD3-class
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((u32)acc[i]), "r"(b[i]));
          asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a[i]) : "r"(b[i]), "r"(a[(i + 1) % ILP]));
          asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i]));
        } else if (MODE == 5) {  // the butterfly's multiply mix: 6 WIDE + 4 IMAD
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((u32)acc[i]), "r"(b[i]));
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((u32)(acc[i] >> 32)), "r"(a[i]));
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((u32)acc[i]), "r"(a[i]));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"((u32)acc[i]));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(a[i]), "r"((u32)(acc[i] >> 32)));
        } else if (MODE == 7) {  // IMAD.WIDE with a true 64-bit accumulate (multiplicand from the neighbouring chain)
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((u32)acc[(i + 1) % ILP]), "r"(b[i]));
        } else if (MODE == 8) {  // IMAD and IADD3 on disjoint register sets: do the two pipes overlap?
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(a[(i + 1) % ILP]));
          asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(c[i]) : "r"(d[i]), "r"(c[(i + 1) % ILP]));
        } else if (MODE == 9) {  // IMAD + two-operand ALU instruction (one register source + immediate)
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(a[(i + 1) % ILP]));
          asm volatile("add.u32 %0, %0, 12345;" : "+r"(c[i]));
        } else if (MODE == 10) {  // two IADD3 per IMAD
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(a[(i + 1) % ILP]));
          asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(c[i]) : "r"(d[i]), "r"(c[(i + 1) % ILP]));
          asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(d[i]) : "r"(c[i]), "r"(d[(i + 1) % ILP]));
        } else if (MODE == 14) {  // IMAD + FFMA: fmaheavy vs fmalite
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(a[(i + 1) % ILP]));
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(g[i]), "f"(f[(i + 1) % ILP]));
        } else if (MODE == 15) {  // FFMA alone
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(g[i]), "f"(f[(i + 1) % ILP]));
        } else if (MODE == 17) {  // two IMAD (same pipe) : additive by construction
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(a[(i + 1) % ILP]));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(c[i]) : "r"(d[i]), "r"(c[(i + 1) % ILP]));
        } else if (MODE == 18) {  // WIDE (accumulate) + two two-source adds (a 64-bit add)
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((u32)acc[(i + 1) % ILP]), "r"(b[i]));
          asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(c[i]) : "r"(d[i]));
          asm volatile("addc.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(d[i]));
        } else if (MODE == 19) {  // IMAD + SEL (predicate from a loop-invariant compare)
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(a[(i + 1) % ILP]));
          asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; selp.u32 %0, %0, %1, p;}" : "+r"(c[i]) : "r"(d[i]), "r"(predsrc));
        } else if (MODE == 6) {  // multiply mix + the butterfly's ALU share (13 ALU-class instructions per 10 multiplies)
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((u32)acc[i]), "r"(b[i]));
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((u32)(acc[i] >> 32)), "r"(a[i]));
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((u32)acc[i]), "r"(a[i]));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"((u32)acc[i]));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(a[i]), "r"((u32)(acc[i] >> 32)));
          asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a[i]) : "r"(b[i]), "r"(a[(i + 1) % ILP]));
          asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i]));
          asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(b[i]) : "r"(a[i]), "r"(b[(i + 1) % ILP]));
          asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
          asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a[i]) : "r"(b[i]), "r"(a[(i + 2) % ILP]));
          asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i]));
        }
      }
    }
  }
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i] + a[i] + b[i] + c[i] + d[i] + (u64)f[i] + (u64)g[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char *name, u64 *out, const u32 *seed) {
  const int iters = 4000, grid = 148 * 4;
  k<MODE><<<grid, 256>>>(out, seed, 10);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(out, seed, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // cycles per "group" (one inner-loop body of the mode) per sub-partition: 4 unrolls x ILP groups per iteration per warp, 8 warps per SMSP
  const double groups_per_smsp = (double)iters * 4 * ILP * 8;
  const double cyc = ms * 1e-3 * 1.965e9;
  printf("%-72s %8.3f ms   %6.2f cycles per group per sub-partition\n", name, ms, cyc / groups_per_smsp);
}

int main() {
  u64 *out; u32 *seed; cudaMalloc(&out, 148 * 4 * 256 * 8); cudaMalloc(&seed, 1024);
  u32 h[256]; for (int i = 0; i < 256; ++i) h[i] = 0x9E3779B9u * (i + 1); cudaMemcpy(seed, h, 1024, cudaMemcpyHostToDevice);
  run<0>("IMAD.WIDE.U32 (group = 1 WIDE)", out, seed);
  run<1>("IMAD lo (group = 1 IMAD)", out, seed);
  run<2>("IMAD.HI (group = 1 IMAD.HI)", out, seed);
  run<3>("IADD3 (group = 1 three-input add)", out, seed);
  run<4>("1 WIDE + 2 ALU (group)", out, seed);
  run<5>("3 WIDE + 2 IMAD (group = half a butterfly's multiplies; model 16)", out, seed);
  run<6>("3 WIDE + 2 IMAD + 6 ALU (group = half a butterfly; model fma 16 / alu 12)", out, seed);
  run<7>("IMAD.WIDE.U32 with 64-bit accumulate (group = 1 WIDE)", out, seed);
  run<8>("1 IMAD + 1 IADD3, disjoint registers (group; 2.0 if the pipes overlap, 4.0 if not)", out, seed);
  run<9>("1 IMAD + 1 IADD (register + immediate) (group)", out, seed);
  run<10>("1 IMAD + 2 IADD3 (group; 4.0 if overlapped, 6.0 if additive)", out, seed);
  run<17>("2 IMAD (group; same pipe: 4.0)", out, seed);
  run<19>("1 IMAD + 1 ISETP + 1 SEL (group)", out, seed);
  run<15>("FFMA (group = 1 FFMA)", out, seed);
  run<14>("1 IMAD + 1 FFMA (group; fmaheavy + fmalite)", out, seed);
  run<18>("1 WIDE (accumulate) + 64-bit add as 2 two-source adds (group)", out, seed);
  printf("%s (cycles assume 1.965 GHz; 8 warps per sub-partition, ILP %d)\n", cudaGetErrorString(cudaDeviceSynchronize()), ILP);
  return 0;
}

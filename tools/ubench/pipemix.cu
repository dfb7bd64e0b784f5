// Pipe-overlap micro-benchmarks for sm_100a (round 2): which instruction mixes of a 64-bit Shoup butterfly overlap, and
// does the FP64 pipe run beside the integer pipes?  Every thread runs ILP independent chains; a "group" is one body of the
// mode (listed in main); the table prints cycles per group per sub-partition for 4 / 8 / 12 / 16 warps per sub-partition.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o pipemix pipemix.cu ; run: ./pipemix
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;

constexpr int ILP = 8;

#define WIDE(i)  asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a[i]), "r"(b[i]))
#define WIDEU(i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a[i]), "r"(ku))
#define MULW(i)  asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc[i]) : "r"(a[i]), "r"(b[i]))
#define IMAD(i)  asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c[i]) : "r"(a[i]), "r"(b[i]))
#define IMADU(i) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c[i]) : "r"(a[i]), "r"(ku))
#define ADD3(i)  asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(d[i]) : "r"(e[i]), "r"(f[i]))
#define ADD2(i)  asm volatile("add.u32 %0, %0, %1;" : "+r"(d[i]) : "r"(e[i]))
#define ADDI(i)  asm volatile("add.u32 %0, %0, 12345;" : "+r"(d[i]))
#define ADD3B(i) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(e[i]) : "r"(f[i]), "r"(d[i]))
#define LOP(i)   asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(d[i]) : "r"(e[i]), "r"(f[i]))
#define SHF(i)   asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(d[i]) : "r"(e[i]))
#define SETSEL(i) asm volatile("{.reg .pred p; setp.lt.s32 p, %0, 0; selp.u32 %0, %1, %2, p;}" : "+r"(d[i]) : "r"(e[i]), "r"(f[i]))
#define DFMA(i)  asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(x[i]) : "d"(y[i]), "d"(z[i]))
#define FFMA(i)  asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(g[i]) : "f"(h[i]), "f"(h[(i + 1) % ILP]))
#define ADD64(i) asm volatile("add.u64 %0, %0, %1;" : "+l"(acc2[i]) : "l"(acc[i]))

template <int MODE> __global__ void __launch_bounds__(128) k(u64 *out, const u32 *seed, int iters, u32 ku) {
  u64 acc[ILP], acc2[ILP]; u32 a[ILP], b[ILP], c[ILP], d[ILP], e[ILP], f[ILP]; double x[ILP], y[ILP], z[ILP]; float g[ILP], h[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    const u32 s = seed[(threadIdx.x * 13 + i * 7) & 255];
    acc[i] = s; acc2[i] = s * 3; a[i] = s | 1; b[i] = (s >> 3) | 3; c[i] = s ^ 0x55; d[i] = s + 9; e[i] = s * 5; f[i] = s * 11;
    x[i] = (double)(s & 1023) * 1e-3; y[i] = 1.0 + (double)(s & 7) * 1e-9; z[i] = 0.5 + (double)(s & 3) * 1e-9; g[i] = (float)(s & 255); h[i] = 1.0f + (float)(s & 3) * 1e-6f;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (MODE == 0) { WIDE(i); }
        else if (MODE == 1) { MULW(i); }
        else if (MODE == 2) { IMAD(i); }
        else if (MODE == 3) { ADD3(i); }
        else if (MODE == 4) { DFMA(i); }
        else if (MODE == 5) { LOP(i); }
        else if (MODE == 6) { SHF(i); }
        else if (MODE == 7) { SETSEL(i); }
        else if (MODE == 8) { FFMA(i); }
        else if (MODE == 10) { IMAD(i); ADD3(i); }
        else if (MODE == 11) { IMAD(i); ADD2(i); }
        else if (MODE == 12) { IMAD(i); ADDI(i); }
        else if (MODE == 13) { IMADU(i); ADDI(i); }
        else if (MODE == 14) { IMAD(i); LOP(i); }
        else if (MODE == 15) { IMAD(i); SHF(i); }
        else if (MODE == 16) { IMAD(i); ADD3(i); ADD3B(i); }
        else if (MODE == 17) { IMAD(i); FFMA(i); }
        else if (MODE == 18) { ADD3(i); FFMA(i); }
        else if (MODE == 20) { WIDE(i); ADD3(i); }
        else if (MODE == 21) { WIDE(i); ADD3(i); ADD3B(i); }
        else if (MODE == 22) { WIDE(i); ADD3(i); ADD3B(i); LOP(i); }
        else if (MODE == 23) { WIDE(i); ADDI(i); ADDI(i); }
        else if (MODE == 24) { WIDEU(i); ADDI(i); ADDI(i); }
        else if (MODE == 25) { MULW(i); ADD3(i); ADD3B(i); }
        else if (MODE == 26) { WIDE(i); ADD64(i); }
        else if (MODE == 30) { DFMA(i); IMAD(i); }
        else if (MODE == 31) { DFMA(i); ADD3(i); }
        else if (MODE == 32) { DFMA(i); WIDE(i); }
        else if (MODE == 33) { DFMA(i); IMAD(i); ADD3(i); }
        else if (MODE == 34) { DFMA(i); DFMA(i); WIDE(i); ADD3(i); ADD3B(i); }
        else if (MODE == 35) { DFMA(i); FFMA(i); }
        else if (MODE >= 40) { }
      }
      if (MODE >= 40) {
#define REP(OP) _Pragma("unroll") for (int i = 0; i < 4; ++i) { OP(i); }
        if (MODE == 40) {  // today's butterfly mix: 6 WIDE + 4 IMAD + 12 ALU, op-major over 4 independent chains
          REP(WIDE) REP(ADD3) REP(WIDE) REP(ADD3B) REP(WIDE) REP(LOP) REP(WIDE) REP(ADD3) REP(WIDE) REP(ADD3B) REP(WIDE) REP(LOP)
          REP(IMAD) REP(ADD3) REP(IMAD) REP(ADD3B) REP(IMAD) REP(LOP) REP(IMAD) REP(ADD3) REP(ADD3B) REP(LOP)
        } else if (MODE == 41) {  // the same multiplies, 8 ALU
          REP(WIDE) REP(ADD3) REP(WIDE) REP(ADD3B) REP(WIDE) REP(LOP) REP(WIDE) REP(ADD3) REP(WIDE) REP(ADD3B) REP(WIDE) REP(LOP)
          REP(IMAD) REP(ADD3) REP(IMAD) REP(ADD3B) REP(IMAD) REP(IMAD)
        } else if (MODE == 42) {  // the multiplies alone
          REP(WIDE) REP(WIDE) REP(WIDE) REP(WIDE) REP(WIDE) REP(WIDE) REP(IMAD) REP(IMAD) REP(IMAD) REP(IMAD)
        } else if (MODE == 43) {  // FP64-quotient butterfly mix: 2 WIDE + 4 IMAD + 11 DFMA + 14 ALU
          REP(WIDE) REP(DFMA) REP(ADD3) REP(DFMA) REP(ADD3B) REP(WIDE) REP(DFMA) REP(LOP) REP(DFMA) REP(ADD3) REP(IMAD) REP(DFMA) REP(ADD3B) REP(DFMA) REP(LOP)
          REP(IMAD) REP(DFMA) REP(ADD3) REP(DFMA) REP(ADD3B) REP(IMAD) REP(DFMA) REP(LOP) REP(DFMA) REP(ADD3) REP(IMAD) REP(DFMA) REP(ADD3B) REP(LOP) REP(ADD3) REP(ADD3B)
        } else if (MODE == 44) {  // 12 ALU alone
          REP(ADD3) REP(ADD3B) REP(LOP) REP(ADD3) REP(ADD3B) REP(LOP) REP(ADD3) REP(ADD3B) REP(LOP) REP(ADD3) REP(ADD3B) REP(LOP)
        } else if (MODE == 45) {  // 12 ALU with one register operand + immediate
          REP(ADDI) REP(ADDI) REP(ADDI) REP(ADDI) REP(ADDI) REP(ADDI) REP(ADDI) REP(ADDI) REP(ADDI) REP(ADDI) REP(ADDI) REP(ADDI)
        } else if (MODE == 46) {  // multiplies + 12 one-register ALU
          REP(WIDE) REP(ADDI) REP(WIDE) REP(ADDI) REP(WIDE) REP(ADDI) REP(WIDE) REP(ADDI) REP(WIDE) REP(ADDI) REP(WIDE) REP(ADDI)
          REP(IMAD) REP(ADDI) REP(IMAD) REP(ADDI) REP(IMAD) REP(ADDI) REP(IMAD) REP(ADDI) REP(ADDI) REP(ADDI)
        } else if (MODE == 47) {  // multiplies + 12 two-register ALU
          REP(WIDE) REP(ADD2) REP(WIDE) REP(ADD2) REP(WIDE) REP(ADD2) REP(WIDE) REP(ADD2) REP(WIDE) REP(ADD2) REP(WIDE) REP(ADD2)
          REP(IMAD) REP(ADD2) REP(IMAD) REP(ADD2) REP(IMAD) REP(ADD2) REP(IMAD) REP(ADD2) REP(ADD2) REP(ADD2)
        }
#undef REP
      }
    }
  }
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i] + acc2[i] + a[i] + b[i] + c[i] + d[i] + e[i] + f[i] + (u64)x[i] + (u64)g[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char *name, u64 *out, const u32 *seed) {
  const int iters = 1500;
  printf("%-78s", name);
  for (int wps = 4; wps <= 16; wps += 4) {  // warps per sub-partition: wps CTAs of 4 warps per SM (registers permitting)
    const int threads = 128;
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<MODE>, threads, 0);
    if (occ < wps) { printf(" %7s", "-"); continue; }
    k<MODE><<<148 * wps, threads>>>(out, seed, 10, 3u);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148 * wps, threads>>>(out, seed, iters, 3u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double groups_per_smsp = (double)iters * 2 * (MODE >= 40 ? 4 : ILP) * wps;  // (the mixes of a butterfly run 4 chains, op-major)
    printf(" %7.2f", ms * 1e-3 * 1.965e9 / groups_per_smsp);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  printf("\n");
}

int main() {
  u64 *out; u32 *seed; cudaMalloc(&out, (size_t)148 * 16 * 128 * 8); cudaMalloc(&seed, 1024);
  u32 h[256]; for (int i = 0; i < 256; ++i) h[i] = 0x9E3779B9u * (i + 1); cudaMemcpy(seed, h, 1024, cudaMemcpyHostToDevice);
  printf("%-78s %7s %7s %7s %7s   (cycles per group per sub-partition at 1.965 GHz; warps per sub-partition)\n", "group", "4", "8", "12", "16");
  run<0>("WIDE (mad.wide.u32, 64-bit accumulate)", out, seed);
  run<1>("MULW (mul.wide.u32, no accumulate)", out, seed);
  run<2>("IMAD (mad.lo)", out, seed);
  run<3>("ADD3 (three-register add)", out, seed);
  run<4>("DFMA", out, seed);
  run<5>("LOP3", out, seed);
  run<6>("SHF", out, seed);
  run<7>("ISETP + SEL", out, seed);
  run<8>("FFMA", out, seed);
  run<10>("IMAD + ADD3", out, seed);
  run<11>("IMAD + ADD (two registers)", out, seed);
  run<12>("IMAD + ADD (register + immediate)", out, seed);
  run<13>("IMAD (kernel-parameter multiplier) + ADD immediate", out, seed);
  run<14>("IMAD + LOP3", out, seed);
  run<15>("IMAD + SHF", out, seed);
  run<16>("IMAD + 2 ADD3", out, seed);
  run<17>("IMAD + FFMA", out, seed);
  run<18>("ADD3 + FFMA", out, seed);
  run<20>("WIDE + ADD3", out, seed);
  run<21>("WIDE + 2 ADD3", out, seed);
  run<22>("WIDE + 2 ADD3 + LOP3", out, seed);
  run<23>("WIDE + 2 ADD immediate", out, seed);
  run<24>("WIDE (kernel-parameter multiplier) + 2 ADD immediate", out, seed);
  run<25>("MULW + 2 ADD3", out, seed);
  run<26>("WIDE + 64-bit add", out, seed);
  run<30>("DFMA + IMAD", out, seed);
  run<31>("DFMA + ADD3", out, seed);
  run<32>("DFMA + WIDE", out, seed);
  run<33>("DFMA + IMAD + ADD3", out, seed);
  run<34>("2 DFMA + WIDE + 2 ADD3", out, seed);
  run<35>("DFMA + FFMA", out, seed);
  run<42>("6 WIDE + 4 IMAD (a butterfly's multiplies)", out, seed);
  run<41>("6 WIDE + 4 IMAD + 8 ALU", out, seed);
  run<40>("6 WIDE + 4 IMAD + 12 ALU (today's butterfly)", out, seed);
  run<44>("12 ALU (three register operands)", out, seed);
  run<45>("12 ALU (one register + immediate)", out, seed);
  run<46>("6 WIDE + 4 IMAD + 12 one-register ALU", out, seed);
  run<47>("6 WIDE + 4 IMAD + 12 two-register ALU", out, seed);
  run<43>("2 WIDE + 4 IMAD + 11 DFMA + 14 ALU (FP64-quotient butterfly)", out, seed);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

// Micro-benchmark (profiling aid, not product code): issue rates of the instructions the blend loop is made of,
// measured on the box as warp-instructions per clock per SM sub-partition.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

#define REP8(X) X X X X X X X X
enum Op { kFFMA, kFFMA2, kFMUL2, kFADD2, kFMNMX, kFSETP_FSEL, kMUFU, kMIX_SCALAR, kMIX_PACKED, kNumOps };
static const char* kNames[kNumOps] = {"ffma", "ffma2", "fmul2", "fadd2", "fmnmx", "fsetp+fsel", "mufu.ex2", "mix_scalar(8ffma+2mufu+4sel)",
                                      "mix_packed(4ffma2+2mufu+4sel)"};

template <int OP>
__global__ void __launch_bounds__(256) bench(float* out, int iters, float seed) {
  float a0 = seed + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float m = 0.999f, c = 0.001f;
  unsigned long long p0, p1, p2, p3, pm, pc;
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1));
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5));
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(pm) : "f"(m));
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(pc) : "f"(c));
  for (int i = 0; i < iters; ++i) {
    if (OP == kFFMA) {
      REP8(asm volatile("fma.rn.f32 %0, %0, %8, %9; fma.rn.f32 %1, %1, %8, %9; fma.rn.f32 %2, %2, %8, %9; fma.rn.f32 %3, %3, %8, %9;"
                        "fma.rn.f32 %4, %4, %8, %9; fma.rn.f32 %5, %5, %8, %9; fma.rn.f32 %6, %6, %8, %9; fma.rn.f32 %7, %7, %8, %9;"
                        : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(m), "f"(c));)
    } else if (OP == kFFMA2) {
      REP8(asm volatile("fma.rn.f32x2 %0, %0, %4, %5; fma.rn.f32x2 %1, %1, %4, %5; fma.rn.f32x2 %2, %2, %4, %5; fma.rn.f32x2 %3, %3, %4, %5;"
                        "fma.rn.f32x2 %0, %0, %4, %5; fma.rn.f32x2 %1, %1, %4, %5; fma.rn.f32x2 %2, %2, %4, %5; fma.rn.f32x2 %3, %3, %4, %5;"
                        : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3) : "l"(pm), "l"(pc));)
    } else if (OP == kFMUL2) {
      REP8(asm volatile("mul.rn.f32x2 %0, %0, %4; mul.rn.f32x2 %1, %1, %4; mul.rn.f32x2 %2, %2, %4; mul.rn.f32x2 %3, %3, %4;"
                        "mul.rn.f32x2 %0, %0, %4; mul.rn.f32x2 %1, %1, %4; mul.rn.f32x2 %2, %2, %4; mul.rn.f32x2 %3, %3, %4;"
                        : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3) : "l"(pm));)
    } else if (OP == kFADD2) {
      REP8(asm volatile("add.rn.f32x2 %0, %0, %4; add.rn.f32x2 %1, %1, %4; add.rn.f32x2 %2, %2, %4; add.rn.f32x2 %3, %3, %4;"
                        "add.rn.f32x2 %0, %0, %4; add.rn.f32x2 %1, %1, %4; add.rn.f32x2 %2, %2, %4; add.rn.f32x2 %3, %3, %4;"
                        : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3) : "l"(pc));)
    } else if (OP == kFMNMX) {
      REP8(asm volatile("min.f32 %0, %0, %8; min.f32 %1, %1, %8; min.f32 %2, %2, %8; min.f32 %3, %3, %8;"
                        "max.f32 %4, %4, %9; max.f32 %5, %5, %9; max.f32 %6, %6, %9; max.f32 %7, %7, %9;"
                        : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(m), "f"(c));)
    } else if (OP == kFSETP_FSEL) {
      REP8(asm volatile("{.reg .pred q0, q1, q2, q3;\n"
                        "setp.lt.f32 q0, %0, %4; setp.lt.f32 q1, %1, %4; setp.lt.f32 q2, %2, %4; setp.lt.f32 q3, %3, %4;\n"
                        "selp.f32 %0, %1, %0, q0; selp.f32 %1, %2, %1, q1; selp.f32 %2, %3, %2, q2; selp.f32 %3, %0, %3, q3;}"
                        : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3) : "f"(m));)
    } else if (OP == kMUFU) {
      REP8(asm volatile("ex2.approx.ftz.f32 %0, %0; ex2.approx.ftz.f32 %1, %1; ex2.approx.ftz.f32 %2, %2; ex2.approx.ftz.f32 %3, %3;"
                        "ex2.approx.ftz.f32 %4, %4; ex2.approx.ftz.f32 %5, %5; ex2.approx.ftz.f32 %6, %6; ex2.approx.ftz.f32 %7, %7;"
                        : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7));)
    } else if (OP == kMIX_SCALAR) {
      REP8(asm volatile("{.reg .pred q0, q1;\n"
                        "fma.rn.f32 %0, %0, %8, %9; fma.rn.f32 %1, %1, %8, %9; fma.rn.f32 %2, %2, %8, %9; fma.rn.f32 %3, %3, %8, %9;\n"
                        "fma.rn.f32 %4, %4, %8, %9; fma.rn.f32 %5, %5, %8, %9; fma.rn.f32 %6, %6, %8, %9; fma.rn.f32 %7, %7, %8, %9;\n"
                        "ex2.approx.ftz.f32 %0, %0; ex2.approx.ftz.f32 %4, %4;\n"
                        "setp.lt.f32 q0, %1, %8; setp.lt.f32 q1, %5, %8; selp.f32 %2, %3, %2, q0; selp.f32 %6, %7, %6, q1;}"
                        : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(m), "f"(c));)
    } else if (OP == kMIX_PACKED) {
      REP8(asm volatile("{.reg .pred q0, q1;\n"
                        "fma.rn.f32x2 %8, %8, %12, %13; fma.rn.f32x2 %9, %9, %12, %13; fma.rn.f32x2 %10, %10, %12, %13; fma.rn.f32x2 %11, %11, %12, %13;\n"
                        "ex2.approx.ftz.f32 %0, %0; ex2.approx.ftz.f32 %4, %4;\n"
                        "setp.lt.f32 q0, %1, %14; setp.lt.f32 q1, %5, %14; selp.f32 %2, %3, %2, q0; selp.f32 %6, %7, %6, q1;}"
                        : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7), "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3)
                        : "l"(pm), "l"(pc), "f"(m));)
    }
  }
  float lo, hi, acc = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p0));
  acc += lo + hi;
  asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p1));
  acc += lo + hi;
  asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p2));
  acc += lo + hi;
  asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p3));
  acc += lo + hi;
  if (acc == 12345.678f) out[0] = acc;
}

template <int OP>
void run(float* out, int sms, double mhz, int instr_per_rep) {
  const int iters = 2000, blocks = sms * 8;
  bench<OP><<<blocks, 256>>>(out, 10, 1.f);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  bench<OP><<<blocks, 256>>>(out, iters, 1.f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  const double warp_instr = (double)blocks * 8 * iters * 8 * instr_per_rep;
  const double clocks = ms * 1e-3 * mhz * 1e6;
  printf("%-34s %8.3f ms  %.3f warp-instr/clk/SMSP (at %.0f MHz nominal)\n", kNames[OP], ms, warp_instr / clocks / sms / 4, mhz);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1000.0;
  float* out;
  cudaMalloc(&out, 4);
  printf("%s, %d SMs, %.0f MHz\n", p.name, p.multiProcessorCount, mhz);
  const int s = p.multiProcessorCount;
  run<kFFMA>(out, s, mhz, 8);
  run<kFFMA2>(out, s, mhz, 8);
  run<kFMUL2>(out, s, mhz, 8);
  run<kFADD2>(out, s, mhz, 8);
  run<kFMNMX>(out, s, mhz, 8);
  run<kFSETP_FSEL>(out, s, mhz, 8);
  run<kMUFU>(out, s, mhz, 8);
  run<kMIX_SCALAR>(out, s, mhz, 14);
  run<kMIX_PACKED>(out, s, mhz, 10);
  printf("last error: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

// Microbenchmark of the B200 FP64 pipe under instruction mixes (developer tool, not product code).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_mix tools/fp64_mix.cu
// Each loop iteration issues 32 independent-chain DFMAs plus K extra instructions of one kind; the
// slope of time vs K gives the cost of that instruction in FP64-pipe cycles.
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
enum { NONE, LDS64, LDS128, RSQ64H, RSQ32, IMAD, F2F_D2F, F2F_F2D, LDS32, DADDDEP, SHFL, I2F64 };

template <int KIND, int K>
__global__ void __launch_bounds__(256) k(double* sink, int iters, double a, double b) {
  __shared__ __align__(16) double sh[512];
  sh[threadIdx.x] = threadIdx.x;
  sh[threadIdx.x + 256] = 1.0;
  __syncthreads();
  double x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-9 + c;
  unsigned sbase = (unsigned)__cvta_generic_to_shared(sh) + (blockIdx.x & 3) * 16;
  int iacc = threadIdx.x;
  float facc = threadIdx.x;
  double dacc = 0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
    }
#pragma unroll
    for (int e = 0; e < K; ++e) {
      if (KIND == LDS64) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sbase + e * 64)); iacc ^= __double2loint(v); }
      if (KIND == LDS32) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sbase + e * 64)); iacc ^= __float_as_int(v); }
      if (KIND == LDS128) { double v, w; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v), "=d"(w) : "r"(sbase + e * 64)); iacc ^= __double2loint(v) ^ __double2hiint(w); }
      if (KIND == RSQ64H) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[e & 7])); iacc ^= __double2hiint(y); }
      if (KIND == RSQ32) { float y; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__int_as_float(iacc | 0x3f000000))); iacc ^= __float_as_int(y); }
      if (KIND == IMAD) { asm volatile("mad.lo.s32 %0, %0, 3, %1;" : "+r"(iacc) : "r"(i)); }
      if (KIND == F2F_D2F) { float y; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(y) : "d"(x[e & 7])); iacc ^= __float_as_int(y); }
      if (KIND == F2F_F2D) { double y; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(y) : "f"(__int_as_float(iacc))); iacc ^= __double2hiint(y); }
      if (KIND == I2F64) { double y; asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(y) : "r"(iacc)); iacc ^= __double2hiint(y); }
      if (KIND == DADDDEP) { dacc += x[e & 7]; }
      if (KIND == SHFL) { int v; asm volatile("shfl.sync.bfly.b32 %0, %1, 1, 31, -1;" : "=r"(v) : "r"(iacc)); iacc ^= v; }
    }
  }
  double s = iacc + facc + dacc;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += x[c];
  if (s == 123.456) sink[0] = s;
}

static float base_ms = 0;
template <int KIND, int K>
void run(const char* name) {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  double* sink; cudaMalloc(&sink, 8);
  const int iters = 40000, threads = 256, blocks = prop.multiProcessorCount * 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<KIND, K><<<blocks, threads>>>(sink, 100, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  k<KIND, K><<<blocks, threads>>>(sink, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  if (KIND == NONE) base_ms = ms;
  // per SMSP: 8 warps, each iteration 32 DFMA → 64 pipe cycles at peak
  double cyc_per_iter_warp = ms * 1e-3 * 1.965e9 / iters / 8.0;   // SMSP cycles per warp-iteration
  double extra = K ? (ms - base_ms) * 1e-3 * 1.965e9 / iters / 8.0 / K : 0;
  printf("%-22s K=%d: %7.2f ms  cycles/warp-iter %.1f (64 = DFMA peak)  extra cycles per added instr %.2f\n", name, K, ms,
         cyc_per_iter_warp, extra);
  cudaFree(sink);
}

int main() {
  run<NONE, 0>("pure DFMA x32");
  run<LDS64, 1>("LDS.64"); run<LDS64, 4>("LDS.64"); run<LDS64, 8>("LDS.64");
  run<LDS32, 4>("LDS.32"); run<LDS128, 4>("LDS.128"); run<LDS128, 8>("LDS.128");
  run<RSQ64H, 1>("MUFU.RSQ64H"); run<RSQ64H, 2>("MUFU.RSQ64H"); run<RSQ64H, 4>("MUFU.RSQ64H");
  run<RSQ32, 2>("MUFU.RSQ f32"); run<RSQ32, 4>("MUFU.RSQ f32");
  run<IMAD, 4>("IMAD"); run<IMAD, 16>("IMAD");
  run<F2F_D2F, 2>("F2F f64->f32"); run<F2F_F2D, 2>("F2F f32->f64"); run<I2F64, 2>("I2F s32->f64");
  run<DADDDEP, 4>("DADD (dependent acc)"); run<SHFL, 4>("SHFL");
  return 0;
}

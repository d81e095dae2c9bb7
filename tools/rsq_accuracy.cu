// Accuracy of the MUFU.RSQ64H seed and of rsqrt_fast (developer tool).
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "../polymer-stats_b200/csrc/chain_math.cuh"
__global__ void k(double* out, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double maxe = 0, maxerr = 0, maxq = 0;
  for (int i = t; i < n; i += gridDim.x * blockDim.x) {
    // x spans many binades: mantissa sweep × exponent sweep
    const double m = 1.0 + (double)(i % 1000003) / 1000003.0 * 3.0;      // [1,4): both exponent parities
    const double x = ldexp(m, (i % 41) - 20);
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double e = fabs(fma(-x, y0 * y0, 1.0));
    maxe = fmax(maxe, e);
    const double y = pmc::rsqrt_fast(x);
    // residual of the refined value in extended form: 1 - x*y*y evaluated with fma error-free pieces
    const double yy = y * y, yy_lo = fma(y, y, -yy);
    const double r = fma(-x, yy, 1.0) - x * yy_lo;
    maxerr = fmax(maxerr, fabs(r) * 0.5);                                  // relative error of y ≈ r/2
    // quadratic refinement only
    const double h = y0 * fma(-x, y0 * y0, 1.0);
    const double yq = fma(h, 0.5, y0);
    const double qq = yq * yq, qq_lo = fma(yq, yq, -qq);
    maxq = fmax(maxq, fabs(fma(-x, qq, 1.0) - x * qq_lo) * 0.5);
  }
  out[3 * t] = maxe; out[3 * t + 1] = maxerr; out[3 * t + 2] = maxq;
}
int main() {
  const int T = 256 * 148; double* d; cudaMalloc(&d, T * 3 * sizeof(double));
  k<<<148, 256>>>(d, 200000000);
  static double h[T * 3]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double a = 0, b = 0, c = 0;
  for (int i = 0; i < T; ++i) { a = fmax(a, h[3 * i]); b = fmax(b, h[3 * i + 1]); c = fmax(c, h[3 * i + 2]); }
  printf("max |1 - x*y0^2| of MUFU.RSQ64H seed = %.3e (2^%.1f)\n", a, log2(a));
  printf("max relative error of rsqrt_fast (cubic step) = %.3e (%.2f ulp)\n", b, b / 1.11e-16);
  printf("max relative error after a quadratic step only = %.3e\n", c);
  return 0;
}

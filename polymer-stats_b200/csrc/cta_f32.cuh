// cta_f32.cuh — the rectangle heads×tails of a single-monomer trial in FP32 ("ΔU … in fp32 with a stated tolerance").
//
// Opt-in (pmc_set_pair_precision, default FP64).  Only the rectangle — idx·(n−1−idx) pairs whose dipoles do not change
// and whose separation changes by the rigid translation D of the tail — is evaluated in single precision; the row
// {idx}×rest (n−1 pairs, both μ_idx and the separation change), the chain state, the running energy, the acceptance test
// and all accumulators stay FP64.  What makes FP32 accurate enough here:
//   * positions are mirrored RELATIVE TO x_idx, rebuilt from the FP64 master copy for every trial, as a PAIR of floats
//     (hi = fl(d), lo = fl(d − hi)): the representation error of an offset is ~4e-15·|d| instead of 6e-8·|d|, and
//     r_ij = (hi_i − hi_j) + (lo_i − lo_j) carries the rounding error of ONE float subtraction, relative to |r_ij| itself —
//     also for two monomers that nearly touch far away from the rotated one (the large terms);
//   * μ_L·r' = μ_L·r − μ_L·D and μ_B·r' = μ_B·r − μ_B·D keep their FP64 structure, every factor carries a relative
//     error of a few 1e-7, and a pair contributes (new − old) with an absolute error of a few 1e-7 of its own magnitude;
//   * the partial sums leave FP32 after kFlush = 32 terms (FP64 accumulators per thread, FP64 reduction);
//   * what remains is r' = r − D, a float difference of floats: its error relative to |r'| is amplified by
//     (|r| + |D|) / |r'| when a trial brings two monomers much closer than they were — then the new term is huge.
// Stated tolerance (include/polymc.h, tests/test_gpu_fp32.py): error ≤ 2e-6·(1 + amplification) of the magnitude of
// each pair term (the sum of the magnitudes of its two parts, μ·μ/r³ and 3(μ·r̂)(μ·r̂)/r³).
// Measured on C2 (4096 chains × 500 trials): 13.2 M updates/s against 7.8 M in FP64; rectangle alone 1.84×
// (tools/rect_bench_f32.cu).  Without the lo parts it would be 16.0 M (rectangle 2.25×) — and wrong by 6e-8·|offset|/|r| per
// term, i.e. useless for long stretched chains; chunk length and accumulation scheme (profiles/r02c_tune_kflush.txt)
// change the speed by < 4 %.
// MUFU.RSQ (rsqrt.approx.f32, 2 ulp) needs no Newton step at this tolerance: 39 FP32 + 2 MUFU per pair against 43 FP64.
// Included by cta_kernels.cuh between its building blocks (CtaView, Team, pair_g, block_sum) and its kernels.
#pragma once

#ifndef PMC_CTA_BUILDING_BLOCKS
#error "include cta_kernels.cuh, which includes this file after its building blocks"
#endif

namespace pmc {

// Mirrors of the staged chain in shared memory, 48 B per monomer.
struct F32View {
  float4* pa;   // hi parts of {x − x_idx, y − y_idx, z − z_idx}, w: e = μ·D_eff — rebuilt every trial
  float4* pb;   // {μx, μy, μz, 0} — follows the FP64 dipoles (written at load and on accept)
  float4* pc;   // lo parts of the offsets, w: 0 — rebuilt every trial
};

__host__ __device__ inline size_t f32_smem_bytes(int n) { return (size_t)n * 3 * sizeof(float4); }

__device__ __forceinline__ F32View carve_f32(unsigned char* base, int n) {
  F32View F;
  F.pa = reinterpret_cast<float4*>(base);
  F.pb = F.pa + n;
  F.pc = F.pb + n;
  return F;
}

__device__ __forceinline__ float rsqrt_f32(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct LaneItemF {
  float x, y, z;     // offset from x_idx, hi part
  float lx, ly, lz;  // … lo part
  float ax, ay, az;  // μ
  float tx, ty, tz;  // −3μ
  float c;           // −3μ·D
};

__device__ __forceinline__ LaneItemF load_lane_item_f32(const F32View& F, int L, float Dx, float Dy, float Dz) {
  const float4 a = F.pa[L], u = F.pb[L], l = F.pc[L];
  LaneItemF it;
  it.x = a.x; it.y = a.y; it.z = a.z;
  it.lx = l.x; it.ly = l.y; it.lz = l.z;
  it.ax = u.x; it.ay = u.y; it.az = u.z;
  it.tx = -3.0f * u.x; it.ty = -3.0f * u.y; it.tz = -3.0f * u.z;
  it.c = fmaf(it.tz, Dz, fmaf(it.ty, Dy, it.tx * Dx));
  return it;
}

// new − old of one pair, the same 43-operation form as rect_pair (cta_kernels.cuh) with MUFU.RSQ for 1/|r|.
__device__ __forceinline__ float rect_pair_f32(const LaneItemF& it, const float4 a, const float4 u, const float4 l,
                                               float Dx, float Dy, float Dz, float acc) {
  const float rx = (it.x - a.x) + (it.lx - l.x), ry = (it.y - a.y) + (it.ly - l.y), rz = (it.z - a.z) + (it.lz - l.z);
  const float mm = fmaf(it.az, u.z, fmaf(it.ay, u.y, it.ax * u.x));
  const float r2 = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
  const float a3 = fmaf(it.tz, rz, fmaf(it.ty, ry, it.tx * rx));
  const float bb = fmaf(u.z, rz, fmaf(u.y, ry, u.x * rx));
  const float qx = rx - Dx, qy = ry - Dy, qz = rz - Dz;
  const float q2 = fmaf(qz, qz, fmaf(qy, qy, qx * qx));
  const float a3n = a3 - it.c;
  const float bn = bb - a.w;
  const float y = rsqrt_f32(r2), yn = rsqrt_f32(q2);
  const float y2 = y * y, yn2 = yn * yn;
  const float t = fmaf(a3 * bb, y2, mm);
  const float tn = fmaf(a3n * bn, yn2, mm);
  acc = fmaf(tn, yn2 * yn, acc);
  return fmaf(-t, y2 * y, acc);
}

// FP32 partial sums are folded into FP64 after this many terms: once a large term is in a float accumulator every later
// addition rounds at 6e-8 of the PARTIAL SUM, so the error grows with the number of terms that follow it
#ifndef PMC_F32_ACC
#define PMC_F32_ACC 0   // 0: chunk sums converted and added in FP64; 1: float-pair two-sum (experiments)
#endif
#ifndef PMC_KFLUSH
#define PMC_KFLUSH 32
#endif
constexpr int kFlush = PMC_KFLUSH;

// One pass over bundles of NL groups of 32 lane items (groups g0, g0+NL, …: `nb` bundles) against the B broadcast
// items; the (bundle, broadcast item) space is split evenly over the team's warps.  Returns this thread's share.
template <class TEAM, int NL>
__device__ __forceinline__ double rect_pass_f32(const F32View& F, int baseA, int A, int g0, int nb, int baseB, int B,
                                                float Dx, float Dy, float Dz) {
  constexpr int W = TEAM::kWarps;
  const int lane = TEAM::lane(), warp = TEAM::warp();
  const float4* __restrict__ pa = F.pa + baseB;
  const float4* __restrict__ pb = F.pb + baseB;
  const float4* __restrict__ pc = F.pc + baseB;
  const int U = nb * B;
  int u = (int)(((long long)U * warp) / W);
  const int u1 = (int)(((long long)U * (warp + 1)) / W);
  double acc = 0.0;
  if (u >= u1) return acc;
  int bundle = u / B;
  int k = u - bundle * B;
  while (u < u1) {
    LaneItemF it[NL];
    bool valid[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      const int l = (g0 + bundle * NL + j) * 32 + lane;
      valid[j] = l < A;
      it[j] = load_lane_item_f32(F, baseA + min(l, A - 1), Dx, Dy, Dz);
    }
    const int kend = min(B, k + (u1 - u));
    u += kend - k;
#if PMC_F32_ACC == 1
    // chunk sums (≤ kFlush terms each, plain FP32) are added into a float PAIR (hi, lo) by an error-free two-sum: six
    // FP32 operations per chunk and lane item, no conversion and no FP64 in the loop; the pair goes to FP64 once per
    // segment
    float hi[NL], lo[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) hi[j] = lo[j] = 0.0f;
#endif
    while (k < kend) {
      const int kc = min(kend, k + kFlush);
      float a[NL];
#pragma unroll
      for (int j = 0; j < NL; ++j) a[j] = 0.0f;
#pragma unroll(NL >= 4 ? 1 : 2)
      for (; k < kc; ++k) {
        const float4 va = pa[k], vu = pb[k], vl = pc[k];
#pragma unroll
        for (int j = 0; j < NL; ++j) a[j] = rect_pair_f32(it[j], va, vu, vl, Dx, Dy, Dz, a[j]);
      }
#if PMC_F32_ACC == 1
#pragma unroll
      for (int j = 0; j < NL; ++j) {   // two-sum: hi + a = s + err exactly
        const float s = hi[j] + a[j];
        const float bb = s - hi[j];
        lo[j] += (hi[j] - (s - bb)) + (a[j] - bb);
        hi[j] = s;
      }
#else
#pragma unroll
      for (int j = 0; j < NL; ++j) acc += valid[j] ? (double)a[j] : 0.0;   // chunk sums leave FP32 here
#endif
    }
#if PMC_F32_ACC == 1
#pragma unroll
    for (int j = 0; j < NL; ++j) acc += valid[j] ? (double)hi[j] + (double)lo[j] : 0.0;
#endif
    k = 0;
    ++bundle;
  }
  return acc;
}

// Σ over the rectangle of (new − old), this thread's share; F.pa / F.pb hold the current trial's mirrors.
template <class TEAM>
__device__ __forceinline__ double rect_sum_f32(const F32View& F, int baseA, int A, int baseB, int B, float Dx, float Dy,
                                               float Dz) {
  const int G = (A + 31) >> 5;   // groups of 32 lane items
  const int n4 = G >> 2;
  double acc = 0.0;
  if (n4) acc += rect_pass_f32<TEAM, 4>(F, baseA, A, 0, n4, baseB, B, Dx, Dy, Dz);
  int g = n4 * 4;
  if (G - g >= 2) {
    acc += rect_pass_f32<TEAM, 2>(F, baseA, A, g, 1, baseB, B, Dx, Dy, Dz);
    g += 2;
  }
  if (G - g >= 1) acc += rect_pass_f32<TEAM, 1>(F, baseA, A, g, 1, baseB, B, Dx, Dy, Dz);
  return acc;
}

// Mirrors of one trial: offsets from x_idx and e = μ·D_eff for every monomer (the lane side ignores e).
template <class TEAM>
__device__ __forceinline__ void build_mirrors_f32(const CtaView& S, const F32View& F, int n, int idx, float ex, float ey,
                                                  float ez) {
  const double xi = S.sx[idx], yi = S.sy[idx], zi = S.sz[idx];
  for (int k = TEAM::ltid(); k < n; k += TEAM::kLocalThreads) {
    const float4 u = F.pb[k];
    const double dx = S.sx[k] - xi, dy = S.sy[k] - yi, dz = S.sz[k] - zi;
    float4 a, l;
    a.x = (float)dx; a.y = (float)dy; a.z = (float)dz;
    l.x = (float)(dx - (double)a.x); l.y = (float)(dy - (double)a.y); l.z = (float)(dz - (double)a.z);
    a.w = fmaf(u.z, ez, fmaf(u.y, ey, u.x * ex));
    l.w = 0.0f;
    F.pa[k] = a;
    F.pc[k] = l;
  }
}

template <class TEAM>
__device__ __forceinline__ void fill_mu_f32(const CtaView& S, const F32View& F, int n) {
  for (int k = TEAM::ltid(); k < n; k += TEAM::kLocalThreads)
    F.pb[k] = make_float4((float)S.mx[k], (float)S.my[k], (float)S.mz[k], 0.0f);
}

// delta_pairs_partial (cta_kernels.cuh) with the rectangle in FP32; the row {idx}×rest stays FP64.  Contains one team
// barrier (mirrors visible).
template <class TEAM>
__device__ __forceinline__ double delta_pairs_partial_f32(const CtaView& S, const F32View& F, int n, double b, int idx,
                                                          double npx, double npy, double npz, double dnx, double dny,
                                                          double dnz) {
  constexpr int T = TEAM::kThreads;
  const int tid = TEAM::tid();
  const double Dx = b * dnx, Dy = b * dny, Dz = b * dnz;
  const double hx = 0.5 * Dx, hy = 0.5 * Dy, hz = 0.5 * Dz;
  const int H = idx, Tl = n - 1 - idx;
  const bool rect = H > 0 && Tl > 0;
  bool lanes_are_heads = true;
  if (rect) {   // lane side = the one that wastes fewer lanes (bundles of up to 4 groups of 32)
    const long long costH = (long long)((H + 31) >> 5) * Tl;
    const long long costT = (long long)((Tl + 31) >> 5) * H;
    lanes_are_heads = costH <= costT;
  }
  const double sgn = lanes_are_heads ? 1.0 : -1.0;
  const float ex = (float)(sgn * Dx), ey = (float)(sgn * Dy), ez = (float)(sgn * Dz);
  const int baseA = lanes_are_heads ? 0 : idx + 1, A = lanes_are_heads ? H : Tl;
  const int baseB = lanes_are_heads ? idx + 1 : 0, B = lanes_are_heads ? Tl : H;
  if (rect) build_mirrors_f32<TEAM>(S, F, n, idx, ex, ey, ez);
  double acc = 0.0;
  {
    const double xi = S.sx[idx], yi = S.sy[idx], zi = S.sz[idx];
    const double ox = S.mx[idx], oy = S.my[idx], oz = S.mz[idx];
    for (int j = tid; j < n; j += T) {
      if (j == idx) continue;
      const double rx = xi - S.sx[j], ry = yi - S.sy[j], rz = zi - S.sz[j];
      const double s = (j < idx) ? 1.0 : -1.0;
      const double ux = S.mx[j], uy = S.my[j], uz = S.mz[j];
      acc += pair_g(npx, npy, npz, ux, uy, uz, fma(s, hx, rx), fma(s, hy, ry), fma(s, hz, rz)) -
             pair_g(ox, oy, oz, ux, uy, uz, rx, ry, rz);
    }
  }
  TEAM::sync();  // mirrors visible
  if (rect) acc += rect_sum_f32<TEAM>(F, baseA, A, baseB, B, ex, ey, ez);
  return acc;
}

template <int T>
__device__ __forceinline__ double cta_delta_pairs_f32(const CtaView& S, const F32View& F, int n, double b, int idx,
                                                      double npx, double npy, double npz, double dnx, double dny,
                                                      double dnz) {
  const double acc = delta_pairs_partial_f32<Team<T / 32, 0>>(S, F, n, b, idx, npx, npy, npz, dnx, dny, dnz);
  return block_sum<T>(acc, S.part, /*trailing_sync=*/false);
}

}  // namespace pmc

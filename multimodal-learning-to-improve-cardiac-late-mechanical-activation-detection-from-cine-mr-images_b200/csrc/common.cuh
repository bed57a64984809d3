// Shared device helpers for the registration-to-strain kernels (sm_100a).
// Math follows SURVEY.md Appendix A; the CPU oracle restates the same formulas
// in oracle/lddmm.py and oracle/strain.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b2lddmm.h"

#define B2_CHECK_LAUNCH()                         \
  do {                                            \
    cudaError_t e__ = cudaGetLastError();         \
    if (e__ != cudaSuccess) return (int)e__;      \
  } while (0)

#define B2_CUDA(call)                             \
  do {                                            \
    cudaError_t e__ = (call);                     \
    if (e__ != cudaSuccess) return (int)e__;      \
  } while (0)

namespace b2 {

constexpr int kMaxGridY = 65535;

__host__ __device__ inline bool is_pow2(int64_t x) { return x > 0 && (x & (x - 1)) == 0; }

// ---------------------------------------------------------------------------
// Deterministic CTA reduction of two float accumulators (loss epilogue): xor-tree inside each warp, one
// partial per warp in `red` (2 * 32 floats of shared memory), xor-tree over the partials in warp 0.
// Fixed order -> bitwise reproducible.  Result valid in thread 0.  Ends with the partials consumed; callers
// synchronise before reusing `red`.
// ---------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ void block_reduce2(float& a, float& b, float* red, int tid) {
  static_assert(NT % 32 == 0 && NT <= 1024, "whole warps");
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((tid & 31) == 0) { red[tid >> 5] = a; red[32 + (tid >> 5)] = b; }
  __syncthreads();
  if (tid < 32) {
    a = tid < NT / 32 ? red[tid] : 0.f;
    b = tid < NT / 32 ? red[32 + tid] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
  }
}

// ---------------------------------------------------------------------------
// Bilinear taps (A.1).  Indices are background-ruled independently; weights are
// never renormalised.  v00=(i,j) v10=(i+1,j) v01=(i,j+1) v11=(i+1,j+1).
// ---------------------------------------------------------------------------
struct Taps {
  int o00, o10, o01, o11;   // flat offsets i*W + j after the background rule
  float a, b;               // fractional parts along rows / cols
  float m00, m10, m01, m11; // 1 = tap contributes, 0 = outside under the zero rule
};

template <int BG>
__device__ __forceinline__ Taps make_taps(float p0, float p1, int H, int W) {
  Taps t;
  float f0 = floorf(p0), f1 = floorf(p1);
  t.a = p0 - f0;
  t.b = p1 - f1;
  int i0 = (int)fminf(fmaxf(f0, -2.0f), (float)(H + 1));
  int j0 = (int)fminf(fmaxf(f1, -2.0f), (float)(W + 1));
  int i1 = i0 + 1, j1 = j0 + 1;
  if (BG == B2_BG_ZERO) {
    float ri0 = (i0 >= 0 && i0 < H) ? 1.f : 0.f, ri1 = (i1 >= 0 && i1 < H) ? 1.f : 0.f;
    float cj0 = (j0 >= 0 && j0 < W) ? 1.f : 0.f, cj1 = (j1 >= 0 && j1 < W) ? 1.f : 0.f;
    t.m00 = ri0 * cj0; t.m10 = ri1 * cj0; t.m01 = ri0 * cj1; t.m11 = ri1 * cj1;
  } else {
    t.m00 = t.m10 = t.m01 = t.m11 = 1.f;
  }
  i0 = min(max(i0, 0), H - 1); i1 = min(max(i1, 0), H - 1);
  j0 = min(max(j0, 0), W - 1); j1 = min(max(j1, 0), W - 1);
  t.o00 = i0 * W + j0; t.o10 = i1 * W + j0; t.o01 = i0 * W + j1; t.o11 = i1 * W + j1;
  return t;
}

// Same taps for the forward gathers.  Under the clamp rule i0 = clamp(floor, 0, H-1), i1 = clamp(floor + 1, 0, H-1):
// clamping floor to [-1, H-1] in float first (NaN -> -1) gives the same two indices with one max and one add+min
// per axis - 5 instructions fewer per gather, bit-identical results.  (The adjoint kernels keep make_taps: there
// the shorter form costs more in register spills than it saves.)
template <int BG>
__device__ __forceinline__ Taps make_taps_fwd(float p0, float p1, int H, int W) {
  if (BG == B2_BG_ZERO) return make_taps<BG>(p0, p1, H, W);
  Taps t;
  const float f0 = floorf(p0), f1 = floorf(p1);
  t.a = p0 - f0;
  t.b = p1 - f1;
  t.m00 = t.m10 = t.m01 = t.m11 = 1.f;
  const int ig = (int)fminf(fmaxf(f0, -1.0f), (float)(H - 1));
  const int jg = (int)fminf(fmaxf(f1, -1.0f), (float)(W - 1));
  const int i0 = max(ig, 0), i1 = min(ig + 1, H - 1);
  const int j0 = max(jg, 0), j1 = min(jg + 1, W - 1);
  t.o00 = i0 * W + j0; t.o10 = i1 * W + j0; t.o01 = i0 * W + j1; t.o11 = i1 * W + j1;
  return t;
}

template <int BG>
__device__ __forceinline__ float tap_sample(const Taps& t, float v00, float v10, float v01, float v11) {
  float oma = 1.f - t.a, omb = 1.f - t.b;
  if (BG == B2_BG_ZERO) { v00 *= t.m00; v10 *= t.m10; v01 *= t.m01; v11 *= t.m11; }
  // same association as the oracle: sum of (wi*wj)*v in tap order 00, 01, 10, 11
  return (((oma * omb) * v00 + (oma * t.b) * v01) + (t.a * omb) * v10) + (t.a * t.b) * v11;
}

// d(sample)/d(p0), d(sample)/d(p1)
template <int BG>
__device__ __forceinline__ void tap_grad(const Taps& t, float v00, float v10, float v01, float v11,
                                         float& g0, float& g1) {
  if (BG == B2_BG_ZERO) { v00 *= t.m00; v10 *= t.m10; v01 *= t.m01; v11 *= t.m11; }
  g0 = (1.f - t.b) * (v10 - v00) + t.b * (v11 - v01);
  g1 = (1.f - t.a) * (v01 - v00) + t.a * (v11 - v10);
}

// ---------------------------------------------------------------------------
// Pre-aggregated two-plane splat (adjoint of the bilinear gather) for threads that walk CONSECUTIVE rows of
// one column with the lanes of a warp along consecutive columns.  For smooth displacements the 2x2 footprints
// of neighbouring pixels overlap: the right column of lane L is the left column of lane L+1, and the bottom row
// of pixel (r, c) is the top row of pixel (r+1, c).  The right-column contributions are handed to lane L+1 by
// warp shuffle and the bottom-row contributions are carried in registers to the thread's next row, so a pixel
// issues ~1 RED per plane instead of 4 (index compares decide: any mismatch - floor crossing, image edge, warp
// edge - falls back to direct atomics, so the sum is always complete).  All 32 lanes must call converged.
// ---------------------------------------------------------------------------
// Fire-and-forget float reduction into GLOBAL memory.  atomicAdd() with an unused result normally becomes RED too, but
// ptxas keeps the returning form (ATOMG: the issuing warp waits for the old value) in every kernel that also contains
// a __threadfence() - the hand-over of the dynamic schedule - which cost the adjoint kernel 12 % before this was
// spelled out.
__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "f"(v) : "memory");
}

struct SplatCarry {
  int idx;       // flat target index of the carried bottom-left contribution, -1 = empty
  float a, b;    // plane 0 / plane 1 values
};

__device__ __forceinline__ void splat_flush(float* __restrict__ G, int plane, SplatCarry& cy) {
  if (cy.idx >= 0) {
    red_add(G + cy.idx, cy.a);
    red_add(G + plane + cy.idx, cy.b);
  }
  cy.idx = -1;
}

template <int BG>
__device__ __forceinline__ void splat2_agg(float* __restrict__ G, int plane, const Taps& t, float g0, float g1,
                                           SplatCarry& cy, int lane) {
  const float oma = 1.f - t.a, omb = 1.f - t.b;
  float w00 = oma * omb, w01 = oma * t.b, w10 = t.a * omb, w11 = t.a * t.b;
  if (BG == B2_BG_ZERO) { w00 *= t.m00; w01 *= t.m01; w10 *= t.m10; w11 *= t.m11; }
  float aT = w00 * g0, bT = w00 * g1;          // (i0, j0)
  float aB = w10 * g0, bB = w10 * g1;          // (i1, j0)
  const float aTr = w01 * g0, bTr = w01 * g1;  // (i0, j1)
  const float aBr = w11 * g0, bBr = w11 * g1;  // (i1, j1)
  constexpr unsigned full = 0xffffffffu;
  const int n01 = __shfl_up_sync(full, t.o01, 1), n11 = __shfl_up_sync(full, t.o11, 1);
  const float naT = __shfl_up_sync(full, aTr, 1), nbT = __shfl_up_sync(full, bTr, 1);
  const float naB = __shfl_up_sync(full, aBr, 1), nbB = __shfl_up_sync(full, bBr, 1);
  const bool take = lane > 0 && n01 == t.o00 && n11 == t.o10;     // lane-1's right column is my left column
  const bool given = (__shfl_down_sync(full, (int)take, 1) != 0) && lane < 31;
  if (take) { aT += naT; bT += nbT; aB += naB; bB += nbB; }
  if (!given) {
    red_add(G + t.o01, aTr); red_add(G + plane + t.o01, bTr);
    red_add(G + t.o11, aBr); red_add(G + plane + t.o11, bBr);
  }
  if (cy.idx == t.o00) { aT += cy.a; bT += cy.b; }
  else splat_flush(G, plane, cy);
  red_add(G + t.o00, aT);
  red_add(G + plane + t.o00, bT);
  cy.idx = t.o10; cy.a = aB; cy.b = bB;
}

// Bilinear samples of the two planes f, f+plane at (p0,p1).  Warp-uniform fast path: when every lane's
// 2x2 footprint lies inside the image, the four taps are base + {0, 1, W, W+1} (immediate offsets from ONE
// address, no clamps); otherwise the whole warp takes the background-ruled path.  Both paths use the same
// weights and tap values, so results are bit-identical.  All 32 lanes must call this converged.
template <int BG, bool FAST = true>
__device__ __forceinline__ void gather2(const float* f, int plane, float p0, float p1, int H, int W,
                                        float& w0, float& w1) {
  const float f0 = floorf(p0), f1 = floorf(p1);
  const int i0 = (int)f0, j0 = (int)f1;            // cvt saturates, NaN -> 0: the slow path handles those
  const bool interior = ((unsigned)i0 < (unsigned)(H - 1)) & ((unsigned)j0 < (unsigned)(W - 1));
  if (FAST && __all_sync(0xffffffffu, interior)) {
    const float a = p0 - f0, b = p1 - f1, oma = 1.f - a, omb = 1.f - b;
    const float* q = f + i0 * W + j0;
    const float c00 = oma * omb, c01 = oma * b, c10 = a * omb, c11 = a * b;
    w0 = ((c00 * q[0] + c01 * q[1]) + c10 * q[W]) + c11 * q[W + 1];
    q += plane;
    w1 = ((c00 * q[0] + c01 * q[1]) + c10 * q[W]) + c11 * q[W + 1];
  } else {
    const Taps t = make_taps_fwd<BG>(p0, p1, H, W);
    w0 = tap_sample<BG>(t, f[t.o00], f[t.o10], f[t.o01], f[t.o11]);
    f += plane;
    w1 = tap_sample<BG>(t, f[t.o00], f[t.o10], f[t.o01], f[t.o11]);
  }
}

template <int BG>
__device__ __forceinline__ float gather1_ldg(const float* __restrict__ f, float p0, float p1, int H, int W) {
  const float f0 = floorf(p0), f1 = floorf(p1);
  const int i0 = (int)f0, j0 = (int)f1;
  const bool interior = ((unsigned)i0 < (unsigned)(H - 1)) & ((unsigned)j0 < (unsigned)(W - 1));
  if (__all_sync(0xffffffffu, interior)) {
    const float a = p0 - f0, b = p1 - f1, oma = 1.f - a, omb = 1.f - b;
    const float* q = f + i0 * W + j0;
    return (((oma * omb) * __ldg(q) + (oma * b) * __ldg(q + 1)) + (a * omb) * __ldg(q + W)) + (a * b) * __ldg(q + W + 1);
  }
  const Taps t = make_taps_fwd<BG>(p0, p1, H, W);
  return tap_sample<BG>(t, __ldg(f + t.o00), __ldg(f + t.o10), __ldg(f + t.o01), __ldg(f + t.o11));
}

// ---------------------------------------------------------------------------
// Finite differences (A.3): central inside, one-sided at the first/last index.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void diff_idx(int k, int n, int& lo, int& hi, float& s) {
  lo = max(k - 1, 0);
  hi = min(k + 1, n - 1);
  s = (k == 0 || k == n - 1) ? 1.f : 0.5f;
}

// Transposed difference operator, gathered form:
// (D^T q)[k] = [k>=1] s(k-1) q[k-1] + [k==n-1] q[k] - [k<=n-2] s(k+1) q[k+1] - [k==0] q[k]
__device__ __forceinline__ float diff_scale(int j, int n) { return (j == 0 || j == n - 1) ? 1.f : 0.5f; }
__device__ __forceinline__ float diffT(float qm, float q0, float qp, int k, int n) {
  float v = 0.f;
  if (k >= 1) v += diff_scale(k - 1, n) * qm;
  if (k == n - 1) v += q0;
  if (k <= n - 2) v -= diff_scale(k + 1, n) * qp;
  if (k == 0) v -= q0;
  return v;
}

// ---------------------------------------------------------------------------
// Sector classification (D7) - integer predicates only.
// table[2k] = round(2^20 sin(theta0 + 2 pi k/n)), table[2k+1] = round(2^20 cos(theta0 + 2 pi k/n)): the boundary
// directions of the slice's sector frame, rotated on the host (b2_sector_table_rotated_host).  `theta0` only
// seeds the search (any value gives the same result); `flip` (counter-clockwise numbering of the reference's
// spl2patchSA mesh, /root/reference/modules/data/utils/DENSE_utils.py:201-204) maps k -> n-1-k.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int classify_sector(long long dr, long long dc, const int32_t* __restrict__ table, int n,
                                               float theta0 = 0.f, bool flip = false) {
  if (dr == 0 && dc == 0) return -1;
  float th = atan2f((float)dr, (float)dc) - theta0;
  th -= 6.283185307179586f * floorf(th * 0.15915494309189535f);
  int k = (int)floorf(th * ((float)n * 0.15915494309189535f));
  k = min(max(k, 0), n - 1);
  for (int it = 0; it < n; ++it) {
    int k1 = (k + 1 == n) ? 0 : k + 1;
    long long lo = (long long)table[2 * k + 1] * dr - (long long)table[2 * k] * dc;
    long long hi = (long long)table[2 * k1 + 1] * dr - (long long)table[2 * k1] * dc;
    if (lo < 0) k = (k == 0) ? n - 1 : k - 1;
    else if (hi >= 0) k = k1;
    else break;
  }
  return flip ? n - 1 - k : k;
}

// Sector frame of slice b as the kernels use it
struct SectorFrame {
  const int32_t* table;   // this slice's (n,2) Q20 boundary table
  float theta0;
  bool flip;
};
__device__ __forceinline__ SectorFrame sector_frame_of(const int32_t* table, long long table_slice_stride,
                                                       const float* theta0, const int32_t* clockwise, long long b) {
  SectorFrame f;
  f.table = table + (size_t)b * (size_t)table_slice_stride;
  f.theta0 = theta0 ? theta0[b] : 0.f;
  f.flip = clockwise ? (clockwise[b] == 0) : false;
  return f;
}

// Centroid as float: double division then cast; image centre for an empty mask.
__device__ __forceinline__ void centroid_from_moments(const long long* mom, int H, int W, float& c0, float& c1) {
  long long cnt = mom[0];
  if (cnt > 0) {
    c0 = (float)((double)mom[1] / (double)cnt);
    c1 = (float)((double)mom[2] / (double)cnt);
  } else {
    c0 = (float)((H - 1) * 0.5);
    c1 = (float)((W - 1) * 0.5);
  }
}

constexpr float kDetEps = 1e-6f;
constexpr float kRad2Eps = 1e-12f;

// Circumferential Green-Lagrange strain at one pixel (A.7).  Returns false when skipped.
struct EccTerms {
  float G00, G01, G10, G11, det, n0, n1, rad2, t0, t1, den, q;
};
__device__ __forceinline__ bool ecc_eval(float d00, float d01, float d10, float d11,
                                         float X0, float X1, float c0, float c1, EccTerms& e, float& ecc) {
  e.G00 = 1.f + d00; e.G01 = d01; e.G10 = d10; e.G11 = 1.f + d11;
  e.det = e.G00 * e.G11 - e.G01 * e.G10;
  e.n0 = X0 - c0; e.n1 = X1 - c1;
  e.rad2 = e.n0 * e.n0 + e.n1 * e.n1;
  float e0 = -e.n1, e1 = e.n0;
  e.t0 = e.G11 * e0 - e.G01 * e1;
  e.t1 = e.G00 * e1 - e.G10 * e0;
  bool valid = (e.rad2 >= kRad2Eps) && (fabsf(e.det) >= kDetEps);
  e.den = valid ? e.rad2 * e.det * e.det : 1.f;
  e.q = (e.t0 * e.t0 + e.t1 * e.t1) / e.den;
  ecc = valid ? 0.5f * (e.q - 1.f) : 0.f;
  return valid;
}

}  // namespace b2

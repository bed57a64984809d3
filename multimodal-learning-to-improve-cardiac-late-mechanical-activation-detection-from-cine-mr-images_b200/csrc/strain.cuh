// Fused Jacobian-stencil strain + sector-binning device code ([SPEC] SURVEY.md A.7/A.8).
// Shared by strain.cu (op level) and shoot.cu (fused epilogue).
#pragma once
#include "common.cuh"

namespace b2 {

constexpr int kMaxSectors = 1024;

// Sector sums are accumulated in Q31.32 fixed point with 64-bit integer shared-memory atomics: integer
// addition is associative, so the strain matrix is bit-identical run to run regardless of the order in
// which threads arrive (float atomics are not).  Quantisation 2^-32 per pixel is far below fp32 resolution.
constexpr float kFixScale = 4294967296.0f;            // 2^32
constexpr double kFixInv = 1.0 / 4294967296.0;
__device__ __forceinline__ unsigned long long ecc_to_fixed(float ecc) {
  const float lim = 1.0e9f;                            // |Ecc| beyond 1e9 is saturated (degenerate Jacobian)
  return (unsigned long long)__float2ll_rn(fminf(fmaxf(ecc, -lim), lim) * kFixScale);
}
__device__ __forceinline__ float fixed_to_mean(unsigned long long s, int cnt) {
  return (float)((double)(long long)s * kFixInv / (double)max(cnt, 1));
}

// Accumulate one (slice, frame) into shared-memory bins.  `u0/u1` point at the
// displacement of this pair, `mask` at its target-frame mask; all threads of the
// CTA call this; bins must be zeroed + synced before and are synced after.
template <int NT>
__device__ __forceinline__ void strain_bin_frame(const float* u0, const float* u1, const float* __restrict__ mask,
                                                 const long long* mom, const int32_t* tab_s, int n_sectors,
                                                 int H, int W, unsigned long long* sums_s, int* cnts_s, int tid,
                                                 float theta0 = 0.f, bool flip = false) {
  const long long cnt = mom[0], sx = mom[1], sy = mom[2];
  float c0, c1;
  centroid_from_moments(mom, H, W, c0, c1);
  const int N = H * W;
  for (int x = tid; x < N; x += NT) {
    if (!(mask[x] > 0.5f)) continue;
    const int r = x / W, c = x - r * W;
    const int k = classify_sector(cnt * r - sx, cnt * c - sy, tab_s, n_sectors, theta0, flip);
    if (k < 0) continue;
    int rlo, rhi, clo, chi; float sr, sc;
    diff_idx(r, H, rlo, rhi, sr);
    diff_idx(c, W, clo, chi, sc);
    const float d00 = sr * (u0[rhi * W + c] - u0[rlo * W + c]);
    const float d10 = sr * (u1[rhi * W + c] - u1[rlo * W + c]);
    const float d01 = sc * (u0[r * W + chi] - u0[r * W + clo]);
    const float d11 = sc * (u1[r * W + chi] - u1[r * W + clo]);
    EccTerms e; float ecc;
    if (!ecc_eval(d00, d01, d10, d11, (float)r + u0[x], (float)c + u1[x], c0, c1, e, ecc)) continue;
    atomicAdd(&sums_s[k], ecc_to_fixed(ecc));
    atomicAdd(&cnts_s[k], 1);
  }
  __syncthreads();
}

// Same accumulation with the member pixels COMPACTED first (fused forward kernel): the myocardium is ~15 % of the
// image, so in the plain loop most warps carry a few member lanes through the whole classify / stencil / strain path.
// Pass 1 collects the member pixel indices (ballot + one shared-memory counter bump per warp) into `list` (N entries of
// shared memory, 16-bit: N <= 65536), pass 2 runs the heavy path on dense warps.  The sums are integer, so the order
// the list happens to have does not change a bit of the result.  `n_s` must be zero and visible on entry.
template <int NT>
__device__ __forceinline__ void strain_bin_frame_compact(const float* u0, const float* u1, const float* __restrict__ mask,
                                                         const long long* mom, const int32_t* tab_s, int n_sectors,
                                                         int H, int W, unsigned long long* sums_s, int* cnts_s, int tid,
                                                         float theta0, bool flip, unsigned short* list, int* n_s) {
  const int N = H * W, lane = tid & 31;
  for (int x = tid; x < N; x += NT) {            // N % NT == 0: whole warps iterate together
    const bool mem = mask[x] > 0.5f;
    const unsigned bal = __ballot_sync(0xffffffffu, mem);
    if (bal) {
      int base = 0;
      if (lane == 0) base = atomicAdd(n_s, __popc(bal));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (mem) list[base + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)x;
    }
  }
  __syncthreads();
  const int n = *n_s;
  const long long cnt = mom[0], sx = mom[1], sy = mom[2];
  float c0, c1;
  centroid_from_moments(mom, H, W, c0, c1);
  for (int i = tid; i < n; i += NT) {
    const int x = list[i];
    const int r = x / W, c = x - r * W;
    const int k = classify_sector(cnt * r - sx, cnt * c - sy, tab_s, n_sectors, theta0, flip);
    if (k < 0) continue;
    int rlo, rhi, clo, chi; float sr, sc;
    diff_idx(r, H, rlo, rhi, sr);
    diff_idx(c, W, clo, chi, sc);
    const float d00 = sr * (u0[rhi * W + c] - u0[rlo * W + c]);
    const float d10 = sr * (u1[rhi * W + c] - u1[rlo * W + c]);
    const float d01 = sc * (u0[r * W + chi] - u0[r * W + clo]);
    const float d11 = sc * (u1[r * W + chi] - u1[r * W + clo]);
    EccTerms e; float ecc;
    if (!ecc_eval(d00, d01, d10, d11, (float)r + u0[x], (float)c + u1[x], c0, c1, e, ecc)) continue;
    atomicAdd(&sums_s[k], ecc_to_fixed(ecc));
    atomicAdd(&cnts_s[k], 1);
  }
  __syncthreads();
}

// ---- adjoint of the reduction (shared by strain_sector_bwd_kernel and the prologue of the fused EPDiff adjoint)
// dL/dEcc of a member pixel of sector k of pair (b, t): (gS[b,k,t] + the edge-padded columns of the last frame) / count
template <int NT>
__device__ __forceinline__ void strain_bwd_weights(const float* __restrict__ gS, const int32_t* __restrict__ counts, int b,
                                                   int t, int T1, int n_sectors, int n_frames, float* gk_s, int tid) {
  for (int k = tid; k < n_sectors; k += NT) {
    const float* row = gS + ((size_t)b * n_sectors + k) * n_frames;
    float g = (t < n_frames) ? row[t] : 0.f;
    if (t == T1 - 1)
      for (int tt = T1; tt < n_frames; ++tt) g += row[tt];
    const int cn = counts[((size_t)b * n_sectors + k) * T1 + t];
    gk_s[k] = g / (float)max(cn, 1);
  }
}

// Accumulates dL/du of one pair into d0 / d1 with float atomics (d zero- or seed-filled by the caller, complete
// before the call; gk_s and tab_s visible to the CTA).  All threads of the CTA call this.
// `list` (optional): N 16-bit entries of shared memory that are free during the call - the member pixels are then
// compacted first and the heavy path runs on dense warps (see strain_bin_frame_compact); N % NT == 0, N <= 65536.
template <int NT>
__device__ __forceinline__ void strain_bwd_frame(const float* u0, const float* u1, const float* __restrict__ mask,
                                                 const long long* mom, const int32_t* tab_s, int n_sectors, int H, int W,
                                                 const float* gk_s, float* d0, float* d1, int tid, float theta0, bool flip,
                                                 unsigned short* list = nullptr) {
  const long long cnt = mom[0], sx = mom[1], sy = mom[2];
  float c0, c1;
  centroid_from_moments(mom, H, W, c0, c1);
  const int N = H * W;
  __shared__ int n_members_s;
  int n_items = N;
  if (list) {
    if (tid == 0) n_members_s = 0;
    __syncthreads();
    const int lane = tid & 31;
    for (int x = tid; x < N; x += NT) {
      const bool mem = mask[x] > 0.5f;
      const unsigned bal = __ballot_sync(0xffffffffu, mem);
      if (bal) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&n_members_s, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (mem) list[base + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)x;
      }
    }
    __syncthreads();
    n_items = n_members_s;
  }
  for (int it = tid; it < n_items; it += NT) {
    const int x = list ? (int)list[it] : it;
    if (!list && !(mask[x] > 0.5f)) continue;
    const int r = x / W, c = x - r * W;
    const int k = classify_sector(cnt * r - sx, cnt * c - sy, tab_s, n_sectors, theta0, flip);
    if (k < 0) continue;
    int rlo, rhi, clo, chi; float sr, sc;
    diff_idx(r, H, rlo, rhi, sr);
    diff_idx(c, W, clo, chi, sc);
    const float d00 = sr * (u0[rhi * W + c] - u0[rlo * W + c]);
    const float d10 = sr * (u1[rhi * W + c] - u1[rlo * W + c]);
    const float d01 = sc * (u0[r * W + chi] - u0[r * W + clo]);
    const float d11 = sc * (u1[r * W + chi] - u1[r * W + clo]);
    EccTerms e; float ecc;
    if (!ecc_eval(d00, d01, d10, d11, (float)r + u0[x], (float)c + u1[x], c0, c1, e, ecc)) continue;
    const float gq = 0.5f * gk_s[k];
    if (gq == 0.f) continue;
    const float e0 = -e.n1, e1 = e.n0;
    const float g_t0 = gq * 2.f * e.t0 / e.den, g_t1 = gq * 2.f * e.t1 / e.den;
    const float g_den = -gq * e.q / e.den;
    const float g_rad2 = g_den * e.det * e.det;
    const float g_det = g_den * e.rad2 * 2.f * e.det;
    float g_G11 = g_t0 * e0 + g_det * e.G00;
    float g_G01 = -g_t0 * e1 - g_det * e.G10;
    float g_G00 = g_t1 * e1 + g_det * e.G11;
    float g_G10 = -g_t1 * e0 - g_det * e.G01;
    const float g_e0 = g_t0 * e.G11 - g_t1 * e.G10;
    const float g_e1 = -g_t0 * e.G01 + g_t1 * e.G00;
    const float g_n0 = 2.f * e.n0 * g_rad2 + g_e1;
    const float g_n1 = 2.f * e.n1 * g_rad2 - g_e0;
    red_add(d0 + x, g_n0);
    red_add(d1 + x, g_n1);
    // G00 = 1 + d0 u0, G10 = d0 u1 (row differences); G01 = d1 u0, G11 = 1 + d1 u1 (col differences)
    red_add(d0 + rhi * W + c, sr * g_G00); red_add(d0 + rlo * W + c, -sr * g_G00);
    red_add(d1 + rhi * W + c, sr * g_G10); red_add(d1 + rlo * W + c, -sr * g_G10);
    red_add(d0 + r * W + chi, sc * g_G01); red_add(d0 + r * W + clo, -sc * g_G01);
    red_add(d1 + r * W + chi, sc * g_G11); red_add(d1 + r * W + clo, -sc * g_G11);
  }
}

// Write column t of S (B,1,K,n_frames) from the bins, with edge-padding of the
// last frame and cropping beyond n_frames (align_n_frames_to semantics).
template <int NT>
__device__ __forceinline__ void strain_store_column(const unsigned long long* sums_s, const int* cnts_s, float* __restrict__ S,
                                                    int32_t* __restrict__ counts, int b, int t, int T1,
                                                    int n_sectors, int n_frames, int tid) {
  for (int k = tid; k < n_sectors; k += NT) {
    const int cn = cnts_s[k];
    const float v = fixed_to_mean(sums_s[k], cn);
    float* row = S + ((size_t)b * n_sectors + k) * n_frames;
    if (t < n_frames) row[t] = v;
    if (t == T1 - 1)
      for (int tt = T1; tt < n_frames; ++tt) row[tt] = v;
    if (counts) counts[((size_t)b * n_sectors + k) * T1 + t] = cn;
  }
}

}  // namespace b2

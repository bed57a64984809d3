// Mask moments, integer-exact sector map, strain-matrix reduction and its adjoint.
// [SPEC] SURVEY.md 8a rows 17-18; sector count / order pinned by
// /root/reference/modules/data/utils/DENSE_utils.py:177-295 and augmentation/affine.py:52-87.
#include <math.h>

#include "strain.cuh"

namespace b2 {

constexpr int kNTS = 256;

__global__ void __launch_bounds__(kNTS)
mask_moments_kernel(const float* __restrict__ mask0, unsigned long long* __restrict__ mom, int H, int W) {
  const int b = blockIdx.y, N = H * W;
  const float* m = mask0 + (size_t)b * N;
  unsigned long long cnt = 0, sx = 0, sy = 0;
  for (int x = blockIdx.x * kNTS + threadIdx.x; x < N; x += gridDim.x * kNTS) {
    if (m[x] > 0.5f) {
      const int r = x / W;
      cnt += 1; sx += (unsigned)r; sy += (unsigned)(x - r * W);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    sx += __shfl_down_sync(0xffffffffu, sx, o);
    sy += __shfl_down_sync(0xffffffffu, sy, o);
  }
  if ((threadIdx.x & 31) == 0 && cnt) {
    atomicAdd(mom + 3 * b + 0, cnt);
    atomicAdd(mom + 3 * b + 1, sx);
    atomicAdd(mom + 3 * b + 2, sy);
  }
}

__global__ void __launch_bounds__(kNTS)
sector_map_kernel(const long long* __restrict__ mom, const b2_sector_frame fr, int64_t b0, int32_t* __restrict__ sector,
                  int H, int W, int n_sectors) {
  extern __shared__ int32_t tab_s[];
  const int b = blockIdx.y, N = H * W;
  const SectorFrame f = sector_frame_of(fr.table, fr.table_slice_stride, fr.theta0, fr.clockwise, b0 + b);
  for (int i = threadIdx.x; i < 2 * n_sectors; i += kNTS) tab_s[i] = f.table[i];
  __syncthreads();
  const long long cnt = mom[3 * b], sx = mom[3 * b + 1], sy = mom[3 * b + 2];
  for (int x = blockIdx.x * kNTS + threadIdx.x; x < N; x += gridDim.x * kNTS) {
    const int r = x / W, c = x - r * W;
    sector[(size_t)b * N + x] = classify_sector(cnt * r - sx, cnt * c - sy, tab_s, n_sectors, f.theta0, f.flip);
  }
}

// one CTA per (slice b, frame t)
__global__ void __launch_bounds__(kNTS)
strain_sector_fwd_kernel(const float* __restrict__ u, const float* __restrict__ tar, const long long* __restrict__ mom,
                         const b2_sector_frame fr, float* __restrict__ S, int32_t* __restrict__ counts,
                         int B, int T1, int H, int W, int n_sectors, int n_frames, long long p0, long long p1) {
  extern __shared__ __align__(8) unsigned long long smem_q[];
  unsigned long long* sums_s = smem_q;                                   // n_sectors x u64
  int32_t* tab_s = reinterpret_cast<int32_t*>(smem_q + n_sectors);       // 2 n_sectors x i32
  int* cnts_s = tab_s + 2 * n_sectors;                                   // n_sectors x i32
  const int tid = threadIdx.x;
  const int N = H * W;
  for (long long pt = p0 + blockIdx.x; pt < p1; pt += gridDim.x) {
    const int b = (int)(pt / T1), t = (int)(pt % T1);
    const SectorFrame f = sector_frame_of(fr.table, fr.table_slice_stride, fr.theta0, fr.clockwise, b);
    __syncthreads();
    for (int i = tid; i < 2 * n_sectors; i += kNTS) tab_s[i] = f.table[i];      // this slice's (rotated) boundaries
    for (int i = tid; i < n_sectors; i += kNTS) { sums_s[i] = 0ull; cnts_s[i] = 0; }
    __syncthreads();
    const float* u0 = u + (size_t)pt * 2 * N;
    strain_bin_frame<kNTS>(u0, u0 + N, tar + (size_t)pt * N, mom + 3 * b, tab_s, n_sectors, H, W, sums_s, cnts_s, tid,
                           f.theta0, f.flip);
    strain_store_column<kNTS>(sums_s, cnts_s, S, counts, b, t, T1, n_sectors, n_frames, tid);
  }
}

// Adjoint: du accumulated with atomics (du zero-filled by the caller).
__global__ void __launch_bounds__(kNTS)
strain_sector_bwd_kernel(const float* __restrict__ gS, const float* __restrict__ u, const float* __restrict__ tar,
                         const long long* __restrict__ mom, const b2_sector_frame fr,
                         const int32_t* __restrict__ counts, float* __restrict__ du,
                         int B, int T1, int H, int W, int n_sectors, int n_frames) {
  extern __shared__ int32_t smem_i[];
  int32_t* tab_s = smem_i;
  float* gk_s = reinterpret_cast<float*>(smem_i + 2 * n_sectors);   // dL/dEcc per member pixel of sector k
  const int tid = threadIdx.x;
  const int N = H * W;
  for (long long pt = blockIdx.x; pt < (long long)B * T1; pt += gridDim.x) {
    const int b = (int)(pt / T1), t = (int)(pt % T1);
    const SectorFrame f = sector_frame_of(fr.table, fr.table_slice_stride, fr.theta0, fr.clockwise, b);
    __syncthreads();
    for (int i = tid; i < 2 * n_sectors; i += kNTS) tab_s[i] = f.table[i];
    strain_bwd_weights<kNTS>(gS, counts, b, t, T1, n_sectors, n_frames, gk_s, tid);
    __syncthreads();
    const float* u0 = u + (size_t)pt * 2 * N;
    float* d0 = du + (size_t)pt * 2 * N;
    strain_bwd_frame<kNTS>(u0, u0 + N, tar + (size_t)pt * N, mom + 3 * b, tab_s, n_sectors, H, W, gk_s, d0, d0 + N, tid,
                           f.theta0, f.flip);
  }
}

static int check_strain(int64_t B, int64_t T1, int64_t H, int64_t W, int n_sectors, int n_frames) {
  if (B <= 0 || T1 <= 0 || H < 2 || W < 2 || H * W > ((int64_t)1 << 30) || B * T1 > ((int64_t)1 << 40)) return B2_E_SHAPE;
  if (n_sectors < 3 || n_sectors > kMaxSectors || n_frames < 1) return B2_E_PARAM;
  return B2_OK;
}

}  // namespace b2

using namespace b2;

extern "C" int b2_sector_table_rotated_host(int n_sectors, double theta0, int32_t* table_host) {
  if (!table_host) return B2_E_NULL;
  if (n_sectors < 3 || n_sectors > kMaxSectors || !(theta0 == theta0) || fabs(theta0) > 1.0e6) return B2_E_PARAM;
  const double q = 1048576.0, two_pi = 6.283185307179586476925286766559;
  for (int k = 0; k < n_sectors; ++k) {
    const double ang = theta0 + two_pi * (double)k / (double)n_sectors;
    table_host[2 * k] = (int32_t)llrint(q * sin(ang));
    table_host[2 * k + 1] = (int32_t)llrint(q * cos(ang));
  }
  return B2_OK;
}

extern "C" int b2_sector_table_host(int n_sectors, int32_t* table_host) {
  return b2_sector_table_rotated_host(n_sectors, 0.0, table_host);
}

extern "C" int b2_mask_moments(const float* mask0, int64_t* moments, int64_t B, int64_t H, int64_t W, void* stream) {
  if (!mask0 || !moments) return B2_E_NULL;
  if (B <= 0 || H < 1 || W < 1 || H * W > ((int64_t)1 << 30)) return B2_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  B2_CUDA(cudaMemsetAsync(moments, 0, sizeof(int64_t) * 3 * (size_t)B, st));
  const int64_t N = H * W;
  int gx = (int)((N + kNTS * 8 - 1) / (kNTS * 8));
  if (gx < 1) gx = 1;
  for (int64_t b0 = 0; b0 < B; b0 += kMaxGridY) {
    const int64_t bn = (B - b0 < kMaxGridY) ? B - b0 : kMaxGridY;
    mask_moments_kernel<<<dim3(gx, (unsigned)bn), kNTS, 0, st>>>(mask0 + (size_t)b0 * N,
                                                                reinterpret_cast<unsigned long long*>(moments) + 3 * b0,
                                                                (int)H, (int)W);
    B2_CHECK_LAUNCH();
  }
  return B2_OK;
}

extern "C" int b2_sector_map_i32(const int64_t* moments, const int32_t* table, int32_t* sector, int64_t B, int64_t H,
                                 int64_t W, int n_sectors, void* stream) {
  const b2_sector_frame fr{table, 0, nullptr, nullptr};
  return b2_sector_map_i32_ex(moments, &fr, sector, B, H, W, n_sectors, stream);
}

extern "C" int b2_sector_map_i32_ex(const int64_t* moments, const b2_sector_frame* frame, int32_t* sector, int64_t B,
                                    int64_t H, int64_t W, int n_sectors, void* stream) {
  if (!moments || !frame || !frame->table || !sector) return B2_E_NULL;
  if (frame->table_slice_stride < 0) return B2_E_PARAM;
  if (B <= 0 || H < 1 || W < 1 || H * W > ((int64_t)1 << 30)) return B2_E_SHAPE;
  if (n_sectors < 3 || n_sectors > kMaxSectors) return B2_E_PARAM;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t N = H * W;
  int gx = (int)((N + kNTS * 4 - 1) / (kNTS * 4));
  if (gx < 1) gx = 1;
  for (int64_t b0 = 0; b0 < B; b0 += kMaxGridY) {
    const int64_t bn = (B - b0 < kMaxGridY) ? B - b0 : kMaxGridY;
    sector_map_kernel<<<dim3(gx, (unsigned)bn), kNTS, sizeof(int32_t) * 2 * n_sectors, st>>>(
        reinterpret_cast<const long long*>(moments) + 3 * b0, *frame, b0, sector + (size_t)b0 * N, (int)H, (int)W, n_sectors);
    B2_CHECK_LAUNCH();
  }
  return B2_OK;
}

extern "C" int b2_strain_sector_fwd(const float* u, const float* tar, const int64_t* moments, const int32_t* table,
                                    float* S, int32_t* counts, int64_t B, int64_t T1, int64_t H, int64_t W,
                                    int n_sectors, int n_frames, void* stream) {
  const b2_sector_frame fr{table, 0, nullptr, nullptr};
  return b2_strain_sector_fwd_ex(u, tar, moments, &fr, S, counts, B, T1, H, W, n_sectors, n_frames, stream);
}

// Pairs [p0, p0 + np) of the batch only (all pointers address the whole batch): the op-level arm of b2_shoot_fwd with
// a pair range (shoot.cu).
namespace b2 {
int strain_sector_fwd_range(const float* u, const float* tar, const int64_t* moments, const b2_sector_frame* frame,
                            float* S, int32_t* counts, int64_t B, int64_t T1, int64_t H, int64_t W, int n_sectors,
                            int n_frames, int64_t p0, int64_t np, cudaStream_t st) {
  if (!u || !tar || !moments || !frame || !frame->table || !S) return B2_E_NULL;
  if (frame->table_slice_stride < 0) return B2_E_PARAM;
  if (int e = check_strain(B, T1, H, W, n_sectors, n_frames)) return e;
  if (p0 < 0 || np <= 0 || p0 + np > B * T1) return B2_E_SHAPE;
  int64_t grid = np;
  if (grid > (1 << 20)) grid = 1 << 20;
  strain_sector_fwd_kernel<<<(unsigned)grid, kNTS, sizeof(int32_t) * 5 * n_sectors, st>>>(
      u, tar, reinterpret_cast<const long long*>(moments), *frame, S, counts, (int)B, (int)T1, (int)H, (int)W,
      n_sectors, n_frames, (long long)p0, (long long)(p0 + np));
  B2_CHECK_LAUNCH();
  return B2_OK;
}
}  // namespace b2

extern "C" int b2_strain_sector_fwd_ex(const float* u, const float* tar, const int64_t* moments,
                                       const b2_sector_frame* frame, float* S, int32_t* counts, int64_t B, int64_t T1,
                                       int64_t H, int64_t W, int n_sectors, int n_frames, void* stream) {
  if (int e = check_strain(B, T1, H, W, n_sectors, n_frames)) return e;
  return strain_sector_fwd_range(u, tar, moments, frame, S, counts, B, T1, H, W, n_sectors, n_frames, 0, B * T1,
                                 (cudaStream_t)stream);
}

extern "C" int b2_strain_sector_bwd(const float* gS, const float* u, const float* tar, const int64_t* moments,
                                    const int32_t* table, const int32_t* counts, float* du, int64_t B, int64_t T1,
                                    int64_t H, int64_t W, int n_sectors, int n_frames, void* stream) {
  const b2_sector_frame fr{table, 0, nullptr, nullptr};
  return b2_strain_sector_bwd_ex(gS, u, tar, moments, &fr, counts, du, B, T1, H, W, n_sectors, n_frames, stream);
}

extern "C" int b2_strain_sector_bwd_ex(const float* gS, const float* u, const float* tar, const int64_t* moments,
                                       const b2_sector_frame* frame, const int32_t* counts, float* du, int64_t B,
                                       int64_t T1, int64_t H, int64_t W, int n_sectors, int n_frames, void* stream) {
  if (!gS || !u || !tar || !moments || !frame || !frame->table || !counts || !du) return B2_E_NULL;
  if (frame->table_slice_stride < 0) return B2_E_PARAM;
  if (int e = check_strain(B, T1, H, W, n_sectors, n_frames)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  B2_CUDA(cudaMemsetAsync(du, 0, sizeof(float) * 2 * (size_t)B * T1 * H * W, st));
  int64_t grid = B * T1;
  if (grid > (1 << 20)) grid = 1 << 20;
  strain_sector_bwd_kernel<<<(unsigned)grid, kNTS, sizeof(int32_t) * 3 * n_sectors, st>>>(
      gS, u, tar, reinterpret_cast<const long long*>(moments), *frame, counts, du, (int)B, (int)T1, (int)H, (int)W,
      n_sectors, n_frames);
  B2_CHECK_LAUNCH();
  return B2_OK;
}

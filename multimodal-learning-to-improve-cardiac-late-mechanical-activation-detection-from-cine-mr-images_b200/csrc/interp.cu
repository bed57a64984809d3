// lagomorph.interp / splat / compose_disp_vel - forward and adjoint kernels.
// Replaces lagomorph_ext.interp_forward / interp_backward.
//
// HBM-bound gather/scatter work: one thread per pixel with lanes along the
// contiguous W axis, so the loads of u / gout and the stores of out are fully
// coalesced 128-byte lines and the 4-tap gathers of a smooth displacement hit
// the same or adjacent lines (L1).  Grid = pixel blocks x batch.
#include "common.cuh"

namespace b2 {

#ifndef B2_OP_THREADS
#define B2_OP_THREADS 128
#endif
constexpr int kThreads = B2_OP_THREADS;
#ifndef B2_PAIRS_PER_THREAD
#define B2_PAIRS_PER_THREAD 1
#endif
// with kPairsPerThread > 1 every thread walks several pairs (stride gridDim.y) in an unrolled loop, which puts
// that many independent load chains in flight per thread
constexpr int kPairsPerThread = B2_PAIRS_PER_THREAD;

// pixel -> (row, col); W is a power of two for every FFT-sized grid, so the common case is a shift
__device__ __forceinline__ void row_col(int x, int W, int wshift, int& r, int& c) {
  if (wshift >= 0) { r = x >> wshift; c = x & (W - 1); }
  else { r = x / W; c = x - r * W; }
}

// CT = compile-time channel count (1 or 2: the path's images/masks and vector fields); 0 = runtime C.
// out = interp(I, u, dt) [+ dt*u when ADD_U: compose_disp_vel]
template <int BG, bool ADD_U, int CT>
__global__ void __launch_bounds__(kThreads)
interp_fwd_kernel(const float* __restrict__ I, const float* __restrict__ u, float* __restrict__ out,
                  int P, int sI, int gI, int su, int C, int H, int W, int wshift, float dt) {
  const int N = H * W;
  const int x = blockIdx.x * kThreads + threadIdx.x;
  if (x >= N) return;
  int r, c;
  row_col(x, W, wshift, r, c);
  const int nc = CT ? CT : C;
#pragma unroll (kPairsPerThread)
  for (int p = blockIdx.y; p < P; p += gridDim.y) {
    const float* up = u + (size_t)(p * su) * 2 * N + x;
    const float u0 = up[0], u1 = up[N];
    const Taps t = make_taps_fwd<BG>((float)r + dt * u0, (float)c + dt * u1, H, W);
    const float* Ic = I + (size_t)((p / gI) * sI) * nc * N;
    float* op = out + (size_t)p * nc * N + x;
#pragma unroll
    for (int ch = 0; ch < (CT ? CT : 1); ++ch) {
      for (int cc = 0; cc < (CT ? 1 : C); ++cc) {
        float v = tap_sample<BG>(t, Ic[t.o00], Ic[t.o10], Ic[t.o01], Ic[t.o11]);
        if (ADD_U) v += dt * ((CT ? ch : cc) == 0 ? u0 : u1);
        *op = v;
        Ic += N;
        op += N;
      }
    }
  }
}

// dI += splat(gout) ; du = dt * sum_c gout_c * grad I_c(x + dt u) [+ dt*gout when ADD_U]
template <int BG, bool ADD_U, bool NEED_DI, bool NEED_DU, int CT>
__global__ void __launch_bounds__(kThreads)
interp_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ I, const float* __restrict__ u,
                  float* __restrict__ dI, float* __restrict__ du,
                  int P, int sI, int gI, int su, int C, int H, int W, int wshift, float dt) {
  const int N = H * W;
  const int x = blockIdx.x * kThreads + threadIdx.x;
  if (x >= N) return;
  int r, c;
  row_col(x, W, wshift, r, c);
  const int nc = CT ? CT : C;
#pragma unroll (kPairsPerThread)
  for (int p = blockIdx.y; p < P; p += gridDim.y) {
    const float* up = u + (size_t)(p * su) * 2 * N + x;
    const float u0 = up[0], u1 = up[N];
    const Taps t = make_taps<BG>((float)r + dt * u0, (float)c + dt * u1, H, W);
    const float oma = 1.f - t.a, omb = 1.f - t.b;
    float w00 = oma * omb, w01 = oma * t.b, w10 = t.a * omb, w11 = t.a * t.b;
    if (BG == B2_BG_ZERO) { w00 *= t.m00; w01 *= t.m01; w10 *= t.m10; w11 *= t.m11; }
    const size_t ioff = (size_t)((p / gI) * sI) * nc * N;
    const float* Ic = I + ioff;
    float* dIc = NEED_DI ? dI + ioff : nullptr;
    const float* gp = gout + (size_t)p * nc * N + x;
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int ch = 0; ch < (CT ? CT : 1); ++ch) {
      for (int cc = 0; cc < (CT ? 1 : C); ++cc) {
        const float g = *gp;
        if (NEED_DI) {
          atomicAdd(dIc + t.o00, w00 * g);
          atomicAdd(dIc + t.o01, w01 * g);
          atomicAdd(dIc + t.o10, w10 * g);
          atomicAdd(dIc + t.o11, w11 * g);
          dIc += N;
        }
        if (NEED_DU) {
          float g0, g1;
          tap_grad<BG>(t, Ic[t.o00], Ic[t.o10], Ic[t.o01], Ic[t.o11], g0, g1);
          a0 += g * g0;
          a1 += g * g1;
          if (ADD_U) { if ((CT ? ch : cc) == 0) a0 += g; else a1 += g; }
        }
        Ic += N;
        gp += N;
      }
    }
    if (NEED_DU) {
      float* dup = du + (size_t)(p * su) * 2 * N + x;
      if (su == 0) {  // broadcast u: reduce over the batch
        atomicAdd(dup, dt * a0);
        atomicAdd(dup + N, dt * a1);
      } else {
        dup[0] = dt * a0;
        dup[N] = dt * a1;
      }
    }
  }
}

template <int BG, int CT>
__global__ void __launch_bounds__(kThreads)
splat_fwd_kernel(const float* __restrict__ J, const float* __restrict__ u, float* __restrict__ out,
                 float* __restrict__ wout, int P, int sJ, int su, int C, int H, int W, int wshift, float dt) {
  const int N = H * W;
  const int x = blockIdx.x * kThreads + threadIdx.x;
  if (x >= N) return;
  int r, c;
  row_col(x, W, wshift, r, c);
  const int nc = CT ? CT : C;
#pragma unroll (kPairsPerThread)
  for (int p = blockIdx.y; p < P; p += gridDim.y) {
    const float* up = u + (size_t)(p * su) * 2 * N + x;
    const Taps t = make_taps<BG>((float)r + dt * up[0], (float)c + dt * up[N], H, W);
    const float oma = 1.f - t.a, omb = 1.f - t.b;
    float w00 = oma * omb, w01 = oma * t.b, w10 = t.a * omb, w11 = t.a * t.b;
    if (BG == B2_BG_ZERO) { w00 *= t.m00; w01 *= t.m01; w10 *= t.m10; w11 *= t.m11; }
    const float* Jc = J + (size_t)(p * sJ) * nc * N + x;
    float* d = out + (size_t)p * nc * N;
#pragma unroll
    for (int ch = 0; ch < (CT ? CT : 1); ++ch) {
      for (int cc = 0; cc < (CT ? 1 : C); ++cc) {
        const float g = *Jc;
        atomicAdd(d + t.o00, w00 * g);
        atomicAdd(d + t.o01, w01 * g);
        atomicAdd(d + t.o10, w10 * g);
        atomicAdd(d + t.o11, w11 * g);
        Jc += N;
        d += N;
      }
    }
    if (wout) {
      float* dw = wout + (size_t)p * N;
      atomicAdd(dw + t.o00, w00);
      atomicAdd(dw + t.o01, w01);
      atomicAdd(dw + t.o10, w10);
      atomicAdd(dw + t.o11, w11);
    }
  }
}

// ------------------------------------------------------------------ TMA-staged tile variant of the forward gather
// One CTA = one (pair, band of kTileRows rows).  The rows of I the band can reach - the band plus a halo of
// kTileHalo rows above and below, full width, i.e. ONE contiguous chunk per channel - are staged in shared memory by
// the TMA engine (cp.async.bulk global -> shared, completion on an mbarrier; UBLKCP in SASS) while the threads
// already fetch their displacements.  The four taps of a pixel are then shared-memory reads; a pixel whose 2x2
// footprint leaves the staged rows (|dt u_0| > halo) falls back to the global gather, so the result is the same
// for any displacement.  Same taps and same arithmetic as interp_fwd_kernel: bit-identical output.
#ifndef B2_INTERP_TMA
#define B2_INTERP_TMA 1
#endif
constexpr int kTileThreads = 256;
constexpr size_t kTileMaxBytes = 48 * 1024;      // per CTA: 4+ CTAs per SM keep enough copies in flight (measured)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "B2_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra B2_DONE;\n"
      "bra B2_WAIT;\n"
      "B2_DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

template <int BG, bool ADD_U, int CT>
__global__ void __launch_bounds__(kTileThreads)
interp_fwd_tile_kernel(const float* __restrict__ I, const float* __restrict__ u, float* __restrict__ out,
                       int P, int sI, int gI, int su, int H, int W, int wshift, float dt, int kTileRows, int kTileHalo) {
  extern __shared__ __align__(128) unsigned char tile_raw[];
  __shared__ __align__(8) uint64_t bar;
  float* tile = reinterpret_cast<float*>(tile_raw);
  const int tid = threadIdx.x, N = H * W;
  const int r0 = blockIdx.x * kTileRows;
  const int rlo = max(r0 - kTileHalo, 0), rhi = min(r0 + kTileRows + kTileHalo, H);
  const int rows = min(kTileRows, H - r0);
  const int tile_floats = (rhi - rlo) * W;            // per channel
  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  uint32_t phase = 0;
  for (int p = blockIdx.y; p < P; p += gridDim.y) {
    const float* Ic = I + (size_t)((p / gI) * sI) * CT * N;
    if (tid == 0) {
      mbar_expect_tx(&bar, (uint32_t)(CT * tile_floats * sizeof(float)));
#pragma unroll
      for (int ch = 0; ch < CT; ++ch)
        bulk_g2s(tile + ch * tile_floats, Ic + (size_t)ch * N + (size_t)rlo * W, (uint32_t)(tile_floats * sizeof(float)), &bar);
    }
    const float* up = u + (size_t)(p * su) * 2 * N + (size_t)r0 * W;
    float* op = out + (size_t)p * CT * N + (size_t)r0 * W;
    const int npix = rows * W;
    // the first displacements are fetched while the bulk copy is in flight
    float u0n = 0.f, u1n = 0.f;
    if (tid < npix) { u0n = __ldg(up + tid); u1n = __ldg(up + N + tid); }
    mbar_wait(&bar, phase);
    phase ^= 1;
    const int lo = rlo * W, hi = rhi * W;
    for (int x = tid; x < npix; x += kTileThreads) {
      const float u0 = u0n, u1 = u1n;
      const int xn = x + kTileThreads;
      if (xn < npix) { u0n = __ldg(up + xn); u1n = __ldg(up + N + xn); }
      int r, c;
      row_col(x, W, wshift, r, c);
      r += r0;
      const Taps t = make_taps_fwd<BG>((float)r + dt * u0, (float)c + dt * u1, H, W);
      const bool inside = t.o00 >= lo && t.o11 < hi;      // o00 / o11 are the smallest / largest of the four offsets
#pragma unroll
      for (int ch = 0; ch < CT; ++ch) {
        float v;
        if (inside) {
          const float* q = tile + ch * tile_floats - lo;
          v = tap_sample<BG>(t, q[t.o00], q[t.o10], q[t.o01], q[t.o11]);
        } else {
          const float* q = Ic + (size_t)ch * N;
          v = tap_sample<BG>(t, __ldg(q + t.o00), __ldg(q + t.o10), __ldg(q + t.o01), __ldg(q + t.o11));
        }
        if (ADD_U) v += dt * (ch == 0 ? u0 : u1);
        op[(size_t)ch * N + x] = v;
      }
    }
    __syncthreads();      // every thread has finished reading the tile before the next pair's copy overwrites it
  }
}

static int log2_or_neg(int64_t W) {
  if (W <= 0 || (W & (W - 1))) return -1;
  int s = 0;
  while ((int64_t(1) << s) < W) ++s;
  return s;
}

static int check_dims(int64_t P, int64_t PI, int64_t Pu, int64_t C, int64_t H, int64_t W) {
  if (P <= 0 || C <= 0 || H < 2 || W < 2) return B2_E_SHAPE;
  if (H * W > (int64_t)1 << 30 || P > (int64_t)1 << 30 || C > 65535) return B2_E_SHAPE;
  if ((PI != 1 && PI != P) || (Pu != 1 && Pu != P)) return B2_E_BCAST;
  return B2_OK;
}

static dim3 pixel_grid(int64_t P, int64_t N) {
  int64_t gy = (P + kPairsPerThread - 1) / kPairsPerThread;
  return dim3((unsigned)((N + kThreads - 1) / kThreads), (unsigned)(gy < kMaxGridY ? gy : kMaxGridY), 1);
}

template <bool ADD_U>
static int launch_interp_fwd(const float* I, const float* u, float* out, int64_t P, int64_t PI, int64_t Pu,
                             int64_t C, int64_t H, int64_t W, float dt, int bg, cudaStream_t st, int gI = 1) {
  const int sI = (PI == P || gI > 1) ? 1 : 0, su = Pu == P ? 1 : 0, ws = log2_or_neg(W);
  // TMA-staged tiles: 1 or 2 channels, rows that are whole 16-byte units, 16-byte aligned image base; band of 32
  // rows + 8 halo rows each side, or 16 + 4 when that keeps the tile within kTileMaxBytes
  int kTileRows = 32, kTileHalo = 8;
  size_t tile_bytes = sizeof(float) * (size_t)C * (kTileRows + 2 * kTileHalo) * W;
  if (tile_bytes > kTileMaxBytes) {
    kTileRows = 16; kTileHalo = 4;
    tile_bytes = sizeof(float) * (size_t)C * (kTileRows + 2 * kTileHalo) * W;
  }
  if (B2_INTERP_TMA && (C == 1 || C == 2) && W % 4 == 0 && H >= kTileRows && tile_bytes <= kTileMaxBytes &&
      reinterpret_cast<uintptr_t>(I) % 16 == 0) {
    const int64_t bands = (H + kTileRows - 1) / kTileRows;
    int64_t gy = P;
    const int64_t cap = 148 * 32 / bands > 1 ? 148 * 32 / bands : 1;     // a few waves of CTAs; each loops over pairs
    if (gy > cap) gy = cap;
    dim3 tgrid((unsigned)bands, (unsigned)gy, 1);
#define B2_LAUNCH_TILE(BGV, CTV)                                                                                      \
  do {                                                                                                                \
    B2_CUDA(cudaFuncSetAttribute(interp_fwd_tile_kernel<BGV, ADD_U, CTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)tile_bytes));                                                                   \
    interp_fwd_tile_kernel<BGV, ADD_U, CTV><<<tgrid, kTileThreads, tile_bytes, st>>>(I, u, out, (int)P, sI, gI, su,    \
                                                                                      (int)H, (int)W, ws, dt,          \
                                                                                      kTileRows, kTileHalo);           \
  } while (0)
    if (bg == B2_BG_CLAMP) {
      if (C == 1) B2_LAUNCH_TILE(B2_BG_CLAMP, 1); else B2_LAUNCH_TILE(B2_BG_CLAMP, 2);
    } else {
      if (C == 1) B2_LAUNCH_TILE(B2_BG_ZERO, 1); else B2_LAUNCH_TILE(B2_BG_ZERO, 2);
    }
#undef B2_LAUNCH_TILE
    B2_CHECK_LAUNCH();
    return B2_OK;
  }
  dim3 grid = pixel_grid(P, H * W);
#define B2_LAUNCH_FWD(BGV, CTV)                                                                       \
  interp_fwd_kernel<BGV, ADD_U, CTV><<<grid, kThreads, 0, st>>>(I, u, out, (int)P, sI, gI, su, (int)C, \
                                                               (int)H, (int)W, ws, dt)
  if (bg == B2_BG_CLAMP) {
    if (C == 1) B2_LAUNCH_FWD(B2_BG_CLAMP, 1);
    else if (C == 2) B2_LAUNCH_FWD(B2_BG_CLAMP, 2);
    else B2_LAUNCH_FWD(B2_BG_CLAMP, 0);
  } else {
    if (C == 1) B2_LAUNCH_FWD(B2_BG_ZERO, 1);
    else if (C == 2) B2_LAUNCH_FWD(B2_BG_ZERO, 2);
    else B2_LAUNCH_FWD(B2_BG_ZERO, 0);
  }
#undef B2_LAUNCH_FWD
  B2_CHECK_LAUNCH();
  return B2_OK;
}

template <bool ADD_U, int BGV, int CTV>
static void launch_bwd_variant(dim3 grid, const float* gout, const float* I, const float* u, float* dI, float* du,
                               int P, int sI, int gI, int su, int C, int H, int W, int ws, float dt,
                               cudaStream_t st) {
  if (dI && du)
    interp_bwd_kernel<BGV, ADD_U, true, true, CTV><<<grid, kThreads, 0, st>>>(gout, I, u, dI, du, P, sI, gI, su, C, H, W, ws, dt);
  else if (dI)
    interp_bwd_kernel<BGV, ADD_U, true, false, CTV><<<grid, kThreads, 0, st>>>(gout, I, u, dI, du, P, sI, gI, su, C, H, W, ws, dt);
  else
    interp_bwd_kernel<BGV, ADD_U, false, true, CTV><<<grid, kThreads, 0, st>>>(gout, I, u, dI, du, P, sI, gI, su, C, H, W, ws, dt);
}

template <bool ADD_U>
static int launch_interp_bwd(const float* gout, const float* I, const float* u, float* dI, float* du,
                             int64_t P, int64_t PI, int64_t Pu, int64_t C, int64_t H, int64_t W,
                             float dt, int bg, cudaStream_t st, int gI = 1) {
  if (!dI && !du) return B2_OK;
  const int64_t N = H * W;
  dim3 grid = pixel_grid(P, N);
  const int sI = (PI == P || gI > 1) ? 1 : 0, su = Pu == P ? 1 : 0, ws = log2_or_neg(W);
  if (dI) B2_CUDA(cudaMemsetAsync(dI, 0, sizeof(float) * (size_t)PI * C * N, st));
  if (du && su == 0) B2_CUDA(cudaMemsetAsync(du, 0, sizeof(float) * 2 * N, st));
#define B2_BWD(BGV, CTV) \
  launch_bwd_variant<ADD_U, BGV, CTV>(grid, gout, I, u, dI, du, (int)P, sI, gI, su, (int)C, (int)H, (int)W, ws, dt, st)
  if (bg == B2_BG_CLAMP) {
    if (C == 1) B2_BWD(B2_BG_CLAMP, 1);
    else if (C == 2) B2_BWD(B2_BG_CLAMP, 2);
    else B2_BWD(B2_BG_CLAMP, 0);
  } else {
    if (C == 1) B2_BWD(B2_BG_ZERO, 1);
    else if (C == 2) B2_BWD(B2_BG_ZERO, 2);
    else B2_BWD(B2_BG_ZERO, 0);
  }
#undef B2_BWD
  B2_CHECK_LAUNCH();
  return B2_OK;
}

}  // namespace b2

using namespace b2;

extern "C" int b2_interp_fwd(const float* I, const float* u, float* out, int64_t P, int64_t PI, int64_t Pu,
                             int64_t C, int64_t H, int64_t W, float dt, int background, void* stream) {
  if (!I || !u || !out) return B2_E_NULL;
  if (int e = check_dims(P, PI, Pu, C, H, W)) return e;
  if (background != B2_BG_CLAMP && background != B2_BG_ZERO) return B2_E_PARAM;
  return launch_interp_fwd<false>(I, u, out, P, PI, Pu, C, H, W, dt, background, (cudaStream_t)stream);
}

extern "C" int b2_interp_bwd(const float* gout, const float* I, const float* u, float* dI, float* du,
                             int64_t P, int64_t PI, int64_t Pu, int64_t C, int64_t H, int64_t W,
                             float dt, int background, void* stream) {
  if (!gout || !I || !u) return B2_E_NULL;
  if (int e = check_dims(P, PI, Pu, C, H, W)) return e;
  if (background != B2_BG_CLAMP && background != B2_BG_ZERO) return B2_E_PARAM;
  return launch_interp_bwd<false>(gout, I, u, dI, du, P, PI, Pu, C, H, W, dt, background, (cudaStream_t)stream);
}

extern "C" int b2_splat_fwd(const float* J, const float* u, float* out, float* wout, int64_t P, int64_t PJ,
                            int64_t Pu, int64_t C, int64_t H, int64_t W, float dt, int background, void* stream) {
  if (!J || !u || !out) return B2_E_NULL;
  if (int e = check_dims(P, PJ, Pu, C, H, W)) return e;
  if (background != B2_BG_CLAMP && background != B2_BG_ZERO) return B2_E_PARAM;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t N = H * W;
  B2_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)P * C * N, st));
  if (wout) B2_CUDA(cudaMemsetAsync(wout, 0, sizeof(float) * (size_t)P * N, st));
  dim3 grid = pixel_grid(P, N);
  const int sJ = PJ == P ? 1 : 0, su = Pu == P ? 1 : 0;
  const int ws = log2_or_neg(W);
#define B2_SPLAT(BGV, CTV) \
  splat_fwd_kernel<BGV, CTV><<<grid, kThreads, 0, st>>>(J, u, out, wout, (int)P, sJ, su, (int)C, (int)H, (int)W, ws, dt)
  if (background == B2_BG_CLAMP) {
    if (C == 1) B2_SPLAT(B2_BG_CLAMP, 1);
    else if (C == 2) B2_SPLAT(B2_BG_CLAMP, 2);
    else B2_SPLAT(B2_BG_CLAMP, 0);
  } else {
    if (C == 1) B2_SPLAT(B2_BG_ZERO, 1);
    else if (C == 2) B2_SPLAT(B2_BG_ZERO, 2);
    else B2_SPLAT(B2_BG_ZERO, 0);
  }
#undef B2_SPLAT
  B2_CHECK_LAUNCH();
  return B2_OK;
}

extern "C" int b2_compose_fwd(const float* u, const float* v, float* out, int64_t P, int64_t H, int64_t W,
                              float dt, int background, void* stream) {
  if (!u || !v || !out) return B2_E_NULL;
  if (int e = check_dims(P, P, P, 2, H, W)) return e;
  if (background != B2_BG_CLAMP && background != B2_BG_ZERO) return B2_E_PARAM;
  return launch_interp_fwd<true>(u, v, out, P, P, P, 2, H, W, dt, background, (cudaStream_t)stream);
}

extern "C" int b2_compose_bwd(const float* gout, const float* u, const float* v, float* du, float* dv,
                              int64_t P, int64_t H, int64_t W, float dt, int background, void* stream) {
  if (!gout || !u || !v) return B2_E_NULL;
  if (int e = check_dims(P, P, P, 2, H, W)) return e;
  if (background != B2_BG_CLAMP && background != B2_BG_ZERO) return B2_E_PARAM;
  // interp(I=u, disp=v): dI -> du, d(disp) -> dv (+ dt*gout from the explicit dt*v term)
  return launch_interp_bwd<true>(gout, u, v, du, dv, P, P, P, 2, H, W, dt, background, (cudaStream_t)stream);
}

// Warp of per-slice images by per-pair displacements: src (B,C,H,W) is shared by the
// T1 frame-pairs of its slice (the reference materialises src.repeat(T1),
// /root/reference/modules/data/__init__.py:109; here it is indexed in-kernel).
extern "C" int b2_warp_fwd(const float* src, const float* u, float* out, int64_t B, int64_t T1, int64_t C,
                           int64_t H, int64_t W, float dt, int background, void* stream) {
  if (!src || !u || !out) return B2_E_NULL;
  if (B <= 0 || T1 <= 0 || T1 > ((int64_t)1 << 30)) return B2_E_SHAPE;
  if (int e = check_dims(B * T1, B * T1, B * T1, C, H, W)) return e;
  if (background != B2_BG_CLAMP && background != B2_BG_ZERO) return B2_E_PARAM;
  return launch_interp_fwd<false>(src, u, out, B * T1, B, B * T1, C, H, W, dt, background, (cudaStream_t)stream, (int)T1);
}

extern "C" int b2_warp_bwd(const float* gout, const float* src, const float* u, float* dsrc, float* du, int64_t B,
                           int64_t T1, int64_t C, int64_t H, int64_t W, float dt, int background, void* stream) {
  if (!gout || !src || !u) return B2_E_NULL;
  if (B <= 0 || T1 <= 0 || T1 > ((int64_t)1 << 30)) return B2_E_SHAPE;
  if (int e = check_dims(B * T1, B * T1, B * T1, C, H, W)) return e;
  if (background != B2_BG_CLAMP && background != B2_BG_ZERO) return B2_E_PARAM;
  return launch_interp_bwd<false>(gout, src, u, dsrc, du, B * T1, B, B * T1, C, H, W, dt, background,
                                  (cudaStream_t)stream, (int)T1);
}

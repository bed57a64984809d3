// lagomorph.interp / splat / compose_disp_vel - forward and adjoint kernels.
// Replaces lagomorph_ext.interp_forward / interp_backward.
//
// HBM-bound gather/scatter work: one thread per pixel with lanes along the
// contiguous W axis, so the loads of u / gout and the stores of out are fully
// coalesced 128-byte lines and the 4-tap gathers of a smooth displacement hit
// the same or adjacent lines (L1).  Grid = pixel blocks x batch.
#include "common.cuh"

namespace b2 {

constexpr int kThreads = 256;

// out = interp(I, u, dt) [+ dt*u when ADD_U: compose_disp_vel]
template <int BG, bool ADD_U>
__global__ void __launch_bounds__(kThreads)
interp_fwd_kernel(const float* __restrict__ I, const float* __restrict__ u, float* __restrict__ out,
                  int P, int sI, int gI, int su, int C, int H, int W, float dt) {
  const int N = H * W;
  const int x = blockIdx.x * kThreads + threadIdx.x;
  if (x >= N) return;
  const int r = x / W, c = x - r * W;
  for (int p = blockIdx.y; p < P; p += gridDim.y) {
    const float* up = u + (size_t)p * su * 2 * N;
    const float u0 = up[x], u1 = up[N + x];
    const Taps t = make_taps<BG>((float)r + dt * u0, (float)c + dt * u1, H, W);
    const float* Ip = I + (size_t)(p / gI) * sI * C * N;
    float* op = out + (size_t)p * C * N;
    for (int ch = 0; ch < C; ++ch) {
      const float* Ic = Ip + (size_t)ch * N;
      float v = tap_sample<BG>(t, Ic[t.o00], Ic[t.o10], Ic[t.o01], Ic[t.o11]);
      if (ADD_U) v += dt * (ch == 0 ? u0 : u1);
      op[(size_t)ch * N + x] = v;
    }
  }
}

// dI += splat(gout) ; du = dt * sum_c gout_c * grad I_c(x + dt u) [+ dt*gout when ADD_U]
template <int BG, bool ADD_U, bool NEED_DI, bool NEED_DU>
__global__ void __launch_bounds__(kThreads)
interp_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ I, const float* __restrict__ u,
                  float* __restrict__ dI, float* __restrict__ du,
                  int P, int sI, int gI, int su, int C, int H, int W, float dt) {
  const int N = H * W;
  const int x = blockIdx.x * kThreads + threadIdx.x;
  if (x >= N) return;
  const int r = x / W, c = x - r * W;
  for (int p = blockIdx.y; p < P; p += gridDim.y) {
    const float* up = u + (size_t)p * su * 2 * N;
    const float u0 = up[x], u1 = up[N + x];
    const Taps t = make_taps<BG>((float)r + dt * u0, (float)c + dt * u1, H, W);
    const float oma = 1.f - t.a, omb = 1.f - t.b;
    float w00 = oma * omb, w01 = oma * t.b, w10 = t.a * omb, w11 = t.a * t.b;
    if (BG == B2_BG_ZERO) { w00 *= t.m00; w01 *= t.m01; w10 *= t.m10; w11 *= t.m11; }
    const float* Ip = I + (size_t)(p / gI) * sI * C * N;
    float* dIp = NEED_DI ? dI + (size_t)(p / gI) * sI * C * N : nullptr;
    const float* gp = gout + (size_t)p * C * N;
    float a0 = 0.f, a1 = 0.f;
    for (int ch = 0; ch < C; ++ch) {
      const float g = gp[(size_t)ch * N + x];
      if (NEED_DI) {
        float* d = dIp + (size_t)ch * N;
        atomicAdd(d + t.o00, w00 * g);
        atomicAdd(d + t.o01, w01 * g);
        atomicAdd(d + t.o10, w10 * g);
        atomicAdd(d + t.o11, w11 * g);
      }
      if (NEED_DU) {
        const float* Ic = Ip + (size_t)ch * N;
        float g0, g1;
        tap_grad<BG>(t, Ic[t.o00], Ic[t.o10], Ic[t.o01], Ic[t.o11], g0, g1);
        a0 += g * g0;
        a1 += g * g1;
        if (ADD_U) { if (ch == 0) a0 += g; else a1 += g; }
      }
    }
    if (NEED_DU) {
      float* dup = du + (size_t)p * su * 2 * N;
      if (su == 0) {  // broadcast u: reduce over the batch
        atomicAdd(dup + x, dt * a0);
        atomicAdd(dup + N + x, dt * a1);
      } else {
        dup[x] = dt * a0;
        dup[N + x] = dt * a1;
      }
    }
  }
}

template <int BG>
__global__ void __launch_bounds__(kThreads)
splat_fwd_kernel(const float* __restrict__ J, const float* __restrict__ u, float* __restrict__ out,
                 float* __restrict__ wout, int P, int sJ, int su, int C, int H, int W, float dt) {
  const int N = H * W;
  const int x = blockIdx.x * kThreads + threadIdx.x;
  if (x >= N) return;
  const int r = x / W, c = x - r * W;
  for (int p = blockIdx.y; p < P; p += gridDim.y) {
    const float* up = u + (size_t)p * su * 2 * N;
    const Taps t = make_taps<BG>((float)r + dt * up[x], (float)c + dt * up[N + x], H, W);
    const float oma = 1.f - t.a, omb = 1.f - t.b;
    float w00 = oma * omb, w01 = oma * t.b, w10 = t.a * omb, w11 = t.a * t.b;
    if (BG == B2_BG_ZERO) { w00 *= t.m00; w01 *= t.m01; w10 *= t.m10; w11 *= t.m11; }
    const float* Jp = J + (size_t)p * sJ * C * N;
    float* op = out + (size_t)p * C * N;
    for (int ch = 0; ch < C; ++ch) {
      const float g = Jp[(size_t)ch * N + x];
      float* d = op + (size_t)ch * N;
      atomicAdd(d + t.o00, w00 * g);
      atomicAdd(d + t.o01, w01 * g);
      atomicAdd(d + t.o10, w10 * g);
      atomicAdd(d + t.o11, w11 * g);
    }
    if (wout) {
      float* d = wout + (size_t)p * N;
      atomicAdd(d + t.o00, w00);
      atomicAdd(d + t.o01, w01);
      atomicAdd(d + t.o10, w10);
      atomicAdd(d + t.o11, w11);
    }
  }
}

static int check_dims(int64_t P, int64_t PI, int64_t Pu, int64_t C, int64_t H, int64_t W) {
  if (P <= 0 || C <= 0 || H < 2 || W < 2) return B2_E_SHAPE;
  if (H * W > (int64_t)1 << 30 || P > (int64_t)1 << 30 || C > 65535) return B2_E_SHAPE;
  if ((PI != 1 && PI != P) || (Pu != 1 && Pu != P)) return B2_E_BCAST;
  return B2_OK;
}

static dim3 pixel_grid(int64_t P, int64_t N) {
  return dim3((unsigned)((N + kThreads - 1) / kThreads), (unsigned)(P < kMaxGridY ? P : kMaxGridY), 1);
}

template <bool ADD_U>
static int launch_interp_fwd(const float* I, const float* u, float* out, int64_t P, int64_t PI, int64_t Pu,
                             int64_t C, int64_t H, int64_t W, float dt, int bg, cudaStream_t st, int gI = 1) {
  dim3 grid = pixel_grid(P, H * W);
  const int sI = (PI == P || gI > 1) ? 1 : 0, su = Pu == P ? 1 : 0;
  if (bg == B2_BG_CLAMP)
    interp_fwd_kernel<B2_BG_CLAMP, ADD_U><<<grid, kThreads, 0, st>>>(I, u, out, (int)P, sI, gI, su, (int)C, (int)H, (int)W, dt);
  else
    interp_fwd_kernel<B2_BG_ZERO, ADD_U><<<grid, kThreads, 0, st>>>(I, u, out, (int)P, sI, gI, su, (int)C, (int)H, (int)W, dt);
  B2_CHECK_LAUNCH();
  return B2_OK;
}

template <bool ADD_U>
static int launch_interp_bwd(const float* gout, const float* I, const float* u, float* dI, float* du,
                             int64_t P, int64_t PI, int64_t Pu, int64_t C, int64_t H, int64_t W,
                             float dt, int bg, cudaStream_t st, int gI = 1) {
  if (!dI && !du) return B2_OK;
  const int64_t N = H * W;
  dim3 grid = pixel_grid(P, N);
  const int sI = (PI == P || gI > 1) ? 1 : 0, su = Pu == P ? 1 : 0;
  if (dI) B2_CUDA(cudaMemsetAsync(dI, 0, sizeof(float) * (size_t)PI * C * N, st));
  if (du && su == 0) B2_CUDA(cudaMemsetAsync(du, 0, sizeof(float) * 2 * N, st));
#define B2_LAUNCH_BWD(BGV, DI, DU)                                                              \
  interp_bwd_kernel<BGV, ADD_U, DI, DU><<<grid, kThreads, 0, st>>>(gout, I, u, dI, du, (int)P, sI, gI, su, \
                                                                  (int)C, (int)H, (int)W, dt)
  if (bg == B2_BG_CLAMP) {
    if (dI && du) B2_LAUNCH_BWD(B2_BG_CLAMP, true, true);
    else if (dI) B2_LAUNCH_BWD(B2_BG_CLAMP, true, false);
    else B2_LAUNCH_BWD(B2_BG_CLAMP, false, true);
  } else {
    if (dI && du) B2_LAUNCH_BWD(B2_BG_ZERO, true, true);
    else if (dI) B2_LAUNCH_BWD(B2_BG_ZERO, true, false);
    else B2_LAUNCH_BWD(B2_BG_ZERO, false, true);
  }
#undef B2_LAUNCH_BWD
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
  if (background == B2_BG_CLAMP)
    splat_fwd_kernel<B2_BG_CLAMP><<<grid, kThreads, 0, st>>>(J, u, out, wout, (int)P, sJ, su, (int)C, (int)H, (int)W, dt);
  else
    splat_fwd_kernel<B2_BG_ZERO><<<grid, kThreads, 0, st>>>(J, u, out, wout, (int)P, sJ, su, (int)C, (int)H, (int)W, dt);
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

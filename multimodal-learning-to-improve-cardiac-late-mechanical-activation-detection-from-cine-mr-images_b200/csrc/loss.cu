// Loss epilogue of the registration path.
// RegistrationReconstructionLoss (/root/reference/modules/loss/registration_losses.py:22-28) is
//   0.5 * mean((tar - Sdef)^2) / sigma^2 + w * sum(v * m) / numel(tar);
// its two full-tensor reductions are produced per frame-pair by the shooting kernels
// (b2_shoot_args.loss_terms) or, op-level, by b2_recon_loss_terms below, and their adjoints are taken
// without seed tensors: b2_warp_sqerr_bwd recomputes Sdef from the taps, the regularisation gradient is
// closed-form inside b2_shoot_bwd_loss (shoot.cu).
#include "common.cuh"

namespace b2 {

constexpr int kLossThreads = 256;

// one CTA per frame-pair, fixed summation order (thread-strided partials, then block_reduce2)
__global__ void __launch_bounds__(kLossThreads)
recon_loss_terms_kernel(const float* __restrict__ sdef, const float* __restrict__ tar, const float* __restrict__ vel,
                        const float* __restrict__ m0, float* __restrict__ terms, int64_t P, int N) {
  __shared__ float red[64];
  const int tid = threadIdx.x;
  for (int64_t p = blockIdx.x; p < P; p += gridDim.x) {
    float sq = 0.f, vm = 0.f;
    if (sdef) {
      const float* s = sdef + (size_t)p * N;
      const float* t = tar + (size_t)p * N;
      for (int i = tid; i < N; i += kLossThreads) {
        const float d = __ldg(t + i) - __ldg(s + i);
        sq += d * d;
      }
    }
    if (vel) {
      const float* v = vel + (size_t)p * 2 * N;
      const float* m = m0 + (size_t)p * 2 * N;
      for (int i = tid; i < N; i += kLossThreads) vm += __ldg(v + i) * __ldg(m + i) + __ldg(v + N + i) * __ldg(m + N + i);
    }
    block_reduce2<kLossThreads>(sq, vm, red, tid);
    if (tid == 0) { terms[2 * p] = sq; terms[2 * p + 1] = vm; }
    __syncthreads();
  }
}

constexpr int kWarpThreads = 128;

// du (+)= g_sq[p] * 2 (Sdef - tar) * grad_taps(src);  dsrc += splat.  One thread per pixel, lanes along W.
template <int BG, bool ACC, bool DSRC>
__global__ void __launch_bounds__(kWarpThreads)
warp_sqerr_bwd_kernel(const float* __restrict__ g_sq, const float* __restrict__ src, const float* __restrict__ tar,
                      const float* __restrict__ u, float* __restrict__ du, float* __restrict__ dsrc, int P, int T1,
                      int H, int W, int src_per_pair, int64_t src_ss, int64_t tar_ss) {
  const int N = H * W;
  const int x = blockIdx.x * kWarpThreads + threadIdx.x;
  if (x >= N) return;
  const int r = x / W, c = x - r * W;
  for (int p = blockIdx.y; p < P; p += gridDim.y) {
    const int b = p / T1, t = p - b * T1;
    const size_t soff = src_per_pair ? (src_ss ? (size_t)b * src_ss + (size_t)t * N : (size_t)p * N)
                                     : (size_t)b * (src_ss ? src_ss : N);
    const float* sp = src + soff;
    const float* tp = tar_ss ? tar + (size_t)b * tar_ss + (size_t)t * N : tar + (size_t)p * N;
    const float* up = u + (size_t)p * 2 * N + x;
    const Taps tp4 = make_taps<BG>((float)r + up[0], (float)c + up[N], H, W);
    const float v00 = __ldg(sp + tp4.o00), v10 = __ldg(sp + tp4.o10), v01 = __ldg(sp + tp4.o01), v11 = __ldg(sp + tp4.o11);
    const float sd = tap_sample<BG>(tp4, v00, v10, v01, v11);
    const float g = 2.f * __ldg(g_sq + p) * (sd - __ldg(tp + x));
    float a0, a1;
    tap_grad<BG>(tp4, v00, v10, v01, v11, a0, a1);
    float* dp = du + (size_t)p * 2 * N + x;
    if (ACC) { dp[0] += g * a0; dp[N] += g * a1; }
    else { dp[0] = g * a0; dp[N] = g * a1; }
    if (DSRC) {
      const float oma = 1.f - tp4.a, omb = 1.f - tp4.b;
      float w00 = oma * omb, w01 = oma * tp4.b, w10 = tp4.a * omb, w11 = tp4.a * tp4.b;
      if (BG == B2_BG_ZERO) { w00 *= tp4.m00; w01 *= tp4.m01; w10 *= tp4.m10; w11 *= tp4.m11; }
      float* d = dsrc + (src_per_pair ? (size_t)p * N : (size_t)b * N);
      atomicAdd(d + tp4.o00, w00 * g); atomicAdd(d + tp4.o01, w01 * g);
      atomicAdd(d + tp4.o10, w10 * g); atomicAdd(d + tp4.o11, w11 * g);
    }
  }
}

template <int BG>
static void launch_warp_sqerr_bwd(dim3 grid, cudaStream_t st, bool acc, const float* g_sq, const float* src,
                                  const float* tar, const float* u, float* du, float* dsrc, int P, int T1, int H,
                                  int W, int spp, int64_t sss, int64_t tss) {
#define B2_WSB(A, D) \
  warp_sqerr_bwd_kernel<BG, A, D><<<grid, kWarpThreads, 0, st>>>(g_sq, src, tar, u, du, dsrc, P, T1, H, W, spp, sss, tss)
  if (acc) { if (dsrc) B2_WSB(true, true); else B2_WSB(true, false); }
  else { if (dsrc) B2_WSB(false, true); else B2_WSB(false, false); }
#undef B2_WSB
}

}  // namespace b2

using namespace b2;

extern "C" int b2_recon_loss_terms(const float* sdef, const float* tar, const float* vel, const float* m0,
                                   float* terms, int64_t P, int64_t H, int64_t W, void* stream) {
  if (!terms) return B2_E_NULL;
  if ((sdef == nullptr) != (tar == nullptr) || (vel == nullptr) != (m0 == nullptr)) return B2_E_NULL;
  if (P <= 0 || H < 1 || W < 1 || H * W > ((int64_t)1 << 30) || P > ((int64_t)1 << 30)) return B2_E_SHAPE;
  const int64_t grid = P < 148 * 8 ? P : 148 * 8;
  recon_loss_terms_kernel<<<(unsigned)grid, kLossThreads, 0, (cudaStream_t)stream>>>(sdef, tar, vel, m0, terms, P,
                                                                                    (int)(H * W));
  B2_CHECK_LAUNCH();
  return B2_OK;
}

extern "C" int b2_warp_sqerr_bwd(const float* g_sq, const float* src, const float* tar, const float* u, float* du,
                                 float* dsrc, int64_t B, int64_t T1, int64_t H, int64_t W, int src_per_pair,
                                 int64_t src_slice_stride, int64_t tar_slice_stride, int background, int accumulate,
                                 void* stream) {
  if (!g_sq || !src || !tar || !u || !du) return B2_E_NULL;
  if (B <= 0 || T1 <= 0 || H < 2 || W < 2) return B2_E_SHAPE;
  const int64_t P = B * T1, N = H * W;
  if (N > ((int64_t)1 << 30) || P > ((int64_t)1 << 30)) return B2_E_SHAPE;
  if (src_slice_stride < 0 || tar_slice_stride < 0) return B2_E_PARAM;
  if (background != B2_BG_CLAMP && background != B2_BG_ZERO) return B2_E_PARAM;
  cudaStream_t st = (cudaStream_t)stream;
  if (dsrc) B2_CUDA(cudaMemsetAsync(dsrc, 0, sizeof(float) * (size_t)(src_per_pair ? P : B) * N, st));
  dim3 grid((unsigned)((N + kWarpThreads - 1) / kWarpThreads), (unsigned)(P < kMaxGridY ? P : kMaxGridY), 1);
  if (background == B2_BG_CLAMP)
    launch_warp_sqerr_bwd<B2_BG_CLAMP>(grid, st, accumulate != 0, g_sq, src, tar, u, du, dsrc, (int)P, (int)T1, (int)H,
                                       (int)W, src_per_pair, src_slice_stride, tar_slice_stride);
  else
    launch_warp_sqerr_bwd<B2_BG_ZERO>(grid, st, accumulate != 0, g_sq, src, tar, u, du, dsrc, (int)P, (int)T1, (int)H,
                                      (int)W, src_per_pair, src_slice_stride, tar_slice_stride);
  B2_CHECK_LAUNCH();
  return B2_OK;
}

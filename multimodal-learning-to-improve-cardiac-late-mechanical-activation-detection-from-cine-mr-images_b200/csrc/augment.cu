// Augmentation step in front of the path, on the device:
//   rotate by a multiple of one sector (360/126 deg) + integer translation of the cine masks, and the matching roll
//   of the strain-matrix rows / TOS curve  (/root/reference/modules/data/augmentation/affine.py:24-87,
//   applied as rotate-then-translate, augmentation/__init__.py:20-21).
// The reference rotates with skimage.transform.rotate(order=0, resize=False, mode='constant'): nearest-neighbour
// lookup through the inverse map  in = M * (col, row, 1).  Index selection is integer work, so the map is evaluated
// in float64 with separately rounded multiplies/adds (no FMA contraction) and C round(), exactly as the CPU oracle
// (oracle/augment.py) does in numpy: bit-exact.
#include "common.cuh"

namespace b2 {

constexpr int kAugThreads = 256;

__global__ void __launch_bounds__(kAugThreads)
augment_volume_kernel(const float* __restrict__ vol, float* __restrict__ out, const double* __restrict__ xform,
                      const int32_t* __restrict__ shift, int T, int H, int W) {
  const int N = H * W;
  const int x = blockIdx.x * kAugThreads + threadIdx.x;
  if (x >= N) return;
  const int b = blockIdx.y;
  const int ro = x / W, co = x - ro * W;
  // undo the circular roll: out[(r + ty) mod H, (c + tx) mod W] = rot[r, c]
  int r = ro, c = co;
  if (shift) {
    r = (ro - shift[2 * b]) % H;      if (r < 0) r += H;
    c = (co - shift[2 * b + 1]) % W;  if (c < 0) c += W;
  }
  const double* m = xform + 6 * (size_t)b;
  const double dc = (double)c, dr = (double)r;
  const double xin = __dadd_rn(__dadd_rn(__dmul_rn(m[0], dc), __dmul_rn(m[1], dr)), m[2]);
  const double yin = __dadd_rn(__dadd_rn(__dmul_rn(m[3], dc), __dmul_rn(m[4], dr)), m[5]);
  const double xr = round(xin), yr = round(yin);          // half away from zero, as C round() in skimage's warp
  const bool ok = xr >= 0.0 && xr <= (double)(W - 1) && yr >= 0.0 && yr <= (double)(H - 1);
  const int src = ok ? (int)yr * W + (int)xr : 0;
  const float* ip = vol + (size_t)b * T * N + src;
  float* op = out + (size_t)b * T * N + x;
  for (int t = 0; t < T; ++t) op[(size_t)t * N] = ok ? __ldg(ip + (size_t)t * N) : 0.f;
}

// out[b, (k + n_b) mod R, :] = S[b, k, :]   (np.roll(S, n, axis=0) per sample)
__global__ void __launch_bounds__(kAugThreads)
roll_rows_kernel(const float* __restrict__ S, float* __restrict__ out, const int32_t* __restrict__ n, int R, int C) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * kAugThreads + threadIdx.x;
  if (i >= R * C) return;
  const int ko = i / C, j = i - ko * C;
  int k = (ko - n[b]) % R;
  if (k < 0) k += R;
  out[(size_t)b * R * C + i] = __ldg(S + (size_t)b * R * C + (size_t)k * C + j);
}

}  // namespace b2

using namespace b2;

extern "C" int b2_augment_volume(const float* vol, float* out, const double* xform, const int32_t* shift, int64_t B,
                                 int64_t T, int64_t H, int64_t W, void* stream) {
  if (!vol || !out || !xform) return B2_E_NULL;
  if (vol == out) return B2_E_PARAM;                       // gather: not in place
  if (B <= 0 || T <= 0 || H < 1 || W < 1 || H * W > ((int64_t)1 << 30) || T > ((int64_t)1 << 20)) return B2_E_SHAPE;
  if (B > kMaxGridY) return B2_E_SHAPE;
  dim3 grid((unsigned)((H * W + kAugThreads - 1) / kAugThreads), (unsigned)B, 1);
  augment_volume_kernel<<<grid, kAugThreads, 0, (cudaStream_t)stream>>>(vol, out, xform, shift, (int)T, (int)H, (int)W);
  B2_CHECK_LAUNCH();
  return B2_OK;
}

extern "C" int b2_roll_rows(const float* S, float* out, const int32_t* n, int64_t B, int64_t R, int64_t C,
                            void* stream) {
  if (!S || !out || !n) return B2_E_NULL;
  if (S == out) return B2_E_PARAM;
  if (B <= 0 || R <= 0 || C <= 0 || R * C > ((int64_t)1 << 30) || B > kMaxGridY) return B2_E_SHAPE;
  dim3 grid((unsigned)((R * C + kAugThreads - 1) / kAugThreads), (unsigned)B, 1);
  roll_rows_kernel<<<grid, kAugThreads, 0, (cudaStream_t)stream>>>(S, out, n, (int)R, (int)C);
  B2_CHECK_LAUNCH();
  return B2_OK;
}

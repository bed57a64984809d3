// Slice regrouping behind the path: the per-pair displacement fields of a batch are
// gathered into one (n_slices, C, F, H, W) tensor per slice, cropped to F frames or zero-padded, replacing the Python
// loop + stack + permute + pad of merge_data_of_same_slice_from_batch
// (/root/reference/modules/trainer/joint_registration_regression_trainer.py:54-120).  Pure data movement: one pass,
// 8 B per element (+ the zero fill of padded frames), 128-bit accesses when the image size allows.
#include "common.cuh"

namespace b2 {

constexpr int kRegroupThreads = 256;

// INVERSE = false: out[slot[p]] = u[p] (forward).  INVERSE = true: u[p] = out[slot[p]], zero for a dropped pair -
// the adjoint of the forward map (every slot is written by at most one pair, so the transpose is a plain gather).
template <typename V, bool INVERSE>
__global__ void __launch_bounds__(kRegroupThreads)
regroup_pairs_kernel(V* __restrict__ u, const int32_t* __restrict__ slot, V* __restrict__ out, int64_t P,
                     int C, int F, int64_t nv /* vector elements per (H,W) plane */) {
  for (int64_t p = blockIdx.y; p < P; p += gridDim.y) {
    const int sl = __ldg(slot + p);
    if (sl < 0 && !INVERSE) continue;                      // pair beyond the F frames kept for its slice
    const int s = sl < 0 ? 0 : sl / F, pos = sl < 0 ? 0 : sl - s * F;
    for (int ch = 0; ch < C; ++ch) {
      V* ip = u + ((size_t)p * C + ch) * nv;
      V* op = out + (((size_t)s * C + ch) * F + pos) * nv;
      for (int64_t i = blockIdx.x * (int64_t)kRegroupThreads + threadIdx.x; i < nv; i += (int64_t)gridDim.x * kRegroupThreads) {
        if (!INVERSE) op[i] = ip[i];
        else if (sl >= 0) ip[i] = op[i];
        else ip[i] = V{};
      }
    }
  }
}

template <bool INVERSE>
static int regroup_launch(float* u, const int32_t* pair_slot, float* out, int64_t P, int64_t n_slices, int64_t F,
                          int64_t C, int64_t H, int64_t W, cudaStream_t st) {
  const int64_t N = H * W;
  const bool vec = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(out)) % 16 == 0);
  const int64_t nv = vec ? N / 4 : N;
  int64_t gx = (nv + kRegroupThreads - 1) / kRegroupThreads;
  if (gx > 64) gx = 64;
  dim3 grid((unsigned)gx, (unsigned)(P < kMaxGridY ? P : kMaxGridY), 1);
  if (vec)
    regroup_pairs_kernel<float4, INVERSE><<<grid, kRegroupThreads, 0, st>>>(reinterpret_cast<float4*>(u), pair_slot,
                                                                            reinterpret_cast<float4*>(out), P, (int)C, (int)F, nv);
  else
    regroup_pairs_kernel<float, INVERSE><<<grid, kRegroupThreads, 0, st>>>(u, pair_slot, out, P, (int)C, (int)F, nv);
  B2_CHECK_LAUNCH();
  return B2_OK;
}

}  // namespace b2

using namespace b2;

extern "C" int b2_regroup_pairs(const float* u, const int32_t* pair_slot, float* out, int64_t P, int64_t n_slices,
                                int64_t F, int64_t C, int64_t H, int64_t W, void* stream) {
  if (!u || !pair_slot || !out) return B2_E_NULL;
  if (P <= 0 || n_slices <= 0 || F <= 0 || C <= 0 || H <= 0 || W <= 0) return B2_E_SHAPE;
  const int64_t N = H * W;
  if (N > ((int64_t)1 << 30) || n_slices * F > ((int64_t)1 << 30) || C > 65535) return B2_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  B2_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n_slices * C * F * N, st));   // padded frames are zero
  return regroup_launch<false>(const_cast<float*>(u), pair_slot, out, P, n_slices, F, C, H, W, st);
}

extern "C" int b2_regroup_pairs_bwd(const float* gout, const int32_t* pair_slot, float* gu, int64_t P, int64_t n_slices,
                                    int64_t F, int64_t C, int64_t H, int64_t W, void* stream) {
  if (!gout || !pair_slot || !gu) return B2_E_NULL;
  if (P <= 0 || n_slices <= 0 || F <= 0 || C <= 0 || H <= 0 || W <= 0) return B2_E_SHAPE;
  if (H * W > ((int64_t)1 << 30) || n_slices * F > ((int64_t)1 << 30) || C > 65535) return B2_E_SHAPE;
  return regroup_launch<true>(gu, pair_slot, const_cast<float*>(gout), P, n_slices, F, C, H, W, (cudaStream_t)stream);
}

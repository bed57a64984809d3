// Lossless narrowing of binary cine masks for the host -> device copy.
// The reference feeds the path fp32 volumes whose values are exactly 0 or 1 (binary myocardium masks,
// /root/reference/README.md:21, modules/data/dataset/joint_dataset.py:61-89).  The host-buffer entry point is PCIe-bound,
// so the masks cross the bus as one byte per pixel: a multi-threaded host pass narrows fp32 -> u8 and CHECKS that every
// value is 0 or 1 (any other value makes the caller fall back to the fp32 copy), and a device kernel widens them back
// into the fp32 staging volume the shooting kernel reads in place.  Bit-identical results by construction.
#include <omp.h>

#include "common.cuh"

namespace b2 {

constexpr int kUnpackThreads = 256;

__global__ void __launch_bounds__(kUnpackThreads)
unpack_u8_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int64_t n4) {
  // 4 pixels per thread: one 32-bit load, one 128-bit store
  const uchar4* in4 = reinterpret_cast<const uchar4*>(in);
  float4* out4 = reinterpret_cast<float4*>(out);
  for (int64_t i = (int64_t)blockIdx.x * kUnpackThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kUnpackThreads) {
    const uchar4 v = __ldg(in4 + i);
    out4[i] = make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w);
  }
}

}  // namespace b2

using namespace b2;

// Host: dst[i] = (uint8_t)src[i]; returns 1 when every src[i] is exactly 0.0f or 1.0f, else 0 (dst then unspecified),
// negative on bad arguments.  `threads` <= 0 uses the OpenMP default (the affinity mask of the calling process).
extern "C" int b2_pack_binary_u8_host(const float* src, uint8_t* dst, int64_t n, int threads) {
  if (!src || !dst) return B2_E_NULL;
  if (n < 0) return B2_E_SHAPE;
  int ok = 1;
  const int nt = threads > 0 ? threads : omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nt) reduction(& : ok)
  for (int64_t blk = 0; blk < (n + 65535) / 65536; ++blk) {
    const int64_t lo = blk * 65536, hi = lo + 65536 < n ? lo + 65536 : n;
    int good = 1;
    for (int64_t i = lo; i < hi; ++i) {
      const float v = src[i];
      good &= (v == 0.0f) | (v == 1.0f);
      dst[i] = (uint8_t)(v != 0.0f);
    }
    ok &= good;
  }
  return ok;
}

// Device: out[i] = (float)in[i], i < n.  n must be a multiple of 4 and both pointers 16-byte aligned.
extern "C" int b2_unpack_u8(const uint8_t* in, float* out, int64_t n, void* stream) {
  if (!in || !out) return B2_E_NULL;
  if (n <= 0 || (n & 3)) return B2_E_SHAPE;
  if ((reinterpret_cast<uintptr_t>(in) & 3) || (reinterpret_cast<uintptr_t>(out) & 15)) return B2_E_PARAM;
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + kUnpackThreads - 1) / kUnpackThreads;
  if (blocks > 148 * 8) blocks = 148 * 8;
  unpack_u8_kernel<<<(unsigned)blocks, kUnpackThreads, 0, (cudaStream_t)stream>>>(in, out, n4);
  B2_CHECK_LAUNCH();
  return B2_OK;
}

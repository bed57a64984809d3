// Lossless narrowing of binary cine masks for the host -> device copy.
// The reference feeds the path fp32 volumes whose values are exactly 0 or 1 (binary myocardium masks,
// /root/reference/README.md:21, modules/data/dataset/joint_dataset.py:61-89).  The host-buffer entry point is PCIe-bound,
// so the masks cross the bus as one byte per pixel: a multi-threaded host pass (own worker pool) narrows fp32 -> u8 and CHECKS that every
// value is 0 or 1 (any other value makes the caller fall back to the fp32 copy), and a device kernel widens them back
// into the fp32 staging volume the shooting kernel reads in place.  Bit-identical results by construction.
#include <atomic>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

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

// One bit per pixel (numpy.packbits order: pixel j of a row sits in byte j / 8 at bit 7 - j % 8).  A thread widens the
// four pixels of one nibble into one 128-bit store; the two threads that share a byte read it through L1.
__global__ void __launch_bounds__(kUnpackThreads)
unpack_bits_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int64_t n4) {
  float4* out4 = reinterpret_cast<float4*>(out);
  for (int64_t i = (int64_t)blockIdx.x * kUnpackThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kUnpackThreads) {
    const unsigned b = __ldg(in + (i >> 1));
    const unsigned nib = (i & 1) ? (b & 15u) : (b >> 4);
    out4[i] = make_float4((float)((nib >> 3) & 1u), (float)((nib >> 2) & 1u), (float)((nib >> 1) & 1u), (float)(nib & 1u));
  }
}

}  // namespace b2

using namespace b2;

namespace {

// Small persistent worker pool for the host pass.  Workers sleep on a condition variable between calls (no spinning:
// several ranks share one host), and blocks are handed out dynamically, so an oversubscribed or descheduled worker
// only means the others take more blocks; the calling thread works too.  Never destroyed (threads are detached).
class PackPool {
 public:
  static PackPool& get() { static PackPool* p = new PackPool; return *p; }

  int pack(const float* src, uint8_t* dst, int64_t n, int threads) {
    std::lock_guard<std::mutex> serial(call_);
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    while ((int)nworkers_ < threads - 1) { std::thread(&PackPool::worker, this, nworkers_).detach(); ++nworkers_; }
    {
      std::lock_guard<std::mutex> lk(m_);
      src_ = src; dst_ = dst; n_ = n;
      next_.store(0); ok_.store(1);
      want_ = threads - 1;                 // workers with index < want_ take part in this job
      active_ = want_;
      ++gen_;
    }
    cv_start_.notify_all();
    run_blocks();
    std::unique_lock<std::mutex> lk(m_);
    cv_done_.wait(lk, [&] { return active_ == 0; });
    return ok_.load();
  }

 private:
  static constexpr int64_t kBlock = 1 << 16;

  void run_blocks() {
    // locals: the byte stores below may alias the members as far as the compiler knows, which blocks vectorisation
    const float* __restrict__ src = src_;
    uint8_t* __restrict__ dst = dst_;
    const int64_t n = n_;
    for (;;) {
      const int64_t lo = next_.fetch_add(1) * kBlock;
      if (lo >= n) break;
      const int64_t hi = lo + kBlock < n ? lo + kBlock : n;
      int good = 1;
      int64_t i = lo;
#if defined(__SSE2__)
      {
        const __m128 zero = _mm_setzero_ps(), one = _mm_set1_ps(1.0f);
        __m128 valid = _mm_castsi128_ps(_mm_set1_epi32(-1));
        const __m128i byte1 = _mm_set1_epi8(1);
        for (; i + 16 <= hi; i += 16) {
          const __m128 a0 = _mm_loadu_ps(src + i), a1 = _mm_loadu_ps(src + i + 4);
          const __m128 a2 = _mm_loadu_ps(src + i + 8), a3 = _mm_loadu_ps(src + i + 12);
          const __m128 z0 = _mm_cmpeq_ps(a0, zero), z1 = _mm_cmpeq_ps(a1, zero);
          const __m128 z2 = _mm_cmpeq_ps(a2, zero), z3 = _mm_cmpeq_ps(a3, zero);
          valid = _mm_and_ps(valid, _mm_and_ps(_mm_and_ps(_mm_or_ps(z0, _mm_cmpeq_ps(a0, one)), _mm_or_ps(z1, _mm_cmpeq_ps(a1, one))),
                                               _mm_and_ps(_mm_or_ps(z2, _mm_cmpeq_ps(a2, one)), _mm_or_ps(z3, _mm_cmpeq_ps(a3, one)))));
          // non-zero -> -1 per lane, narrowed with signed saturation to bytes, then & 1
          const __m128i n01 = _mm_packs_epi32(_mm_castps_si128(_mm_cmpneq_ps(a0, zero)), _mm_castps_si128(_mm_cmpneq_ps(a1, zero)));
          const __m128i n23 = _mm_packs_epi32(_mm_castps_si128(_mm_cmpneq_ps(a2, zero)), _mm_castps_si128(_mm_cmpneq_ps(a3, zero)));
          _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), _mm_and_si128(_mm_packs_epi16(n01, n23), byte1));
        }
        good = _mm_movemask_ps(valid) == 0xF;
      }
#endif
      for (; i < hi; ++i) {
        const float v = src[i];
        good &= (v == 0.0f) | (v == 1.0f);
        dst[i] = (uint8_t)(v != 0.0f);
      }
      if (!good) ok_.store(0);
    }
  }

  void worker(size_t index) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_start_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if ((int)index >= want_) continue;   // not part of this job
      }
      run_blocks();
      std::lock_guard<std::mutex> lk(m_);
      if (--active_ == 0) cv_done_.notify_one();
    }
  }

  std::mutex call_, m_;
  std::condition_variable cv_start_, cv_done_;
  size_t nworkers_ = 0;
  uint64_t gen_ = 0;
  int want_ = 0, active_ = 0;
  const float* src_ = nullptr;
  uint8_t* dst_ = nullptr;
  int64_t n_ = 0;
  std::atomic<int64_t> next_{0};
  std::atomic<int> ok_{1};
};

}  // namespace

// Host: dst[i] = (uint8_t)src[i]; returns 1 when every src[i] is exactly 0.0f or 1.0f, else 0 (dst then unspecified),
// negative on bad arguments.  `threads` <= 0 uses std::thread::hardware_concurrency() capped at 16.
extern "C" int b2_pack_binary_u8_host(const float* src, uint8_t* dst, int64_t n, int threads) {
  if (!src || !dst) return B2_E_NULL;
  if (n < 0) return B2_E_SHAPE;
  if (threads <= 0) {
    threads = (int)std::thread::hardware_concurrency();
    if (threads > 16) threads = 16;
  }
  return PackPool::get().pack(src, dst, n, threads);
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

// Device: out[i] = bit i of the numpy.packbits stream `in` (most significant bit first), i < n.  n must be a multiple
// of 8 (whole bytes) and out 16-byte aligned.
extern "C" int b2_unpack_bits(const uint8_t* in, float* out, int64_t n, void* stream) {
  if (!in || !out) return B2_E_NULL;
  if (n <= 0 || (n & 7)) return B2_E_SHAPE;
  if (reinterpret_cast<uintptr_t>(out) & 15) return B2_E_PARAM;
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + kUnpackThreads - 1) / kUnpackThreads;
  if (blocks > 148 * 8) blocks = 148 * 8;
  unpack_bits_kernel<<<(unsigned)blocks, kUnpackThreads, 0, (cudaStream_t)stream>>>(in, out, n4);
  B2_CHECK_LAUNCH();
  return B2_OK;
}

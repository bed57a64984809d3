// lagomorph.FluidMetric.flat / .sharp as shared-memory FFT kernels.
//
// Path A (H*W*8 bytes fit in one SM's shared memory, up to 128x128): ONE persistent
// kernel, one CTA per field at a time: load both components -> complex pack in
// smem -> row FFT -> column FFT -> symbol multiply -> column IFFT -> row IFFT ->
// store.  HBM traffic is the compulsory 16*N bytes per field; the spectrum never
// leaves the SM (the reference does 2 cuFFT launches + a pointwise kernel with
// the spectrum round-tripping through HBM).
//
// Path B (up to 256x256 / rectangular): three kernels with a complex scratch in
// global memory: row FFTs, then column FFT + multiply + column IFFT on a
// column-cell block together with its mirror block, then row IFFTs.
#include "fft.cuh"

namespace b2 {

// ------------------------------------------------------------------ path A
template <int H, int W, int NT, bool INVERSE>
__global__ void __launch_bounds__(NT)
fluid_smem_kernel(const float* __restrict__ f, float* __restrict__ out, int P, FluidParams fp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using FS = FluidSmem<H, W>;
  constexpr int LD = FS::LD, N = H * W;
  float2 *z, *twH, *twW, *csH, *csW;
  FS::carve(smem_raw, z, twH, twW, csH, csW);
  const int tid = threadIdx.x;
  FS::init_luts(twH, twW, csH, csW, tid, NT);
  for (int p = blockIdx.x; p < P; p += gridDim.x) {
    const float* f0 = f + (size_t)p * 2 * N;
    const float* f1 = f0 + N;
    __syncthreads();  // previous iteration's stores done / LUTs visible
    for (int i = tid; i < N; i += NT) {
      const int r = i / W, c = i % W;
      z[r * LD + c] = make_float2(f0[i], f1[i]);
    }
    __syncthreads();
    fluid_smem<H, W, INVERSE, NT>(z, twH, twW, csH, csW, fp, tid);
    float* o0 = out + (size_t)p * 2 * N;
    float* o1 = o0 + N;
    for (int i = tid; i < N; i += NT) {
      const int r = i / W, c = i % W;
      const float2 v = z[r * LD + c];
      o0[i] = v.x;
      o1[i] = v.y;
    }
  }
}

template <int H, int W, int NT>
static int launch_smem(const float* f, float* out, int64_t P, FluidParams fp, int inverse, cudaStream_t st) {
  using FS = FluidSmem<H, W>;
  int dev = 0, sms = 148;
  B2_CUDA(cudaGetDevice(&dev));
  B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = FS::bytes;
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm * NT > 2048) per_sm = 2048 / NT;
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > P) grid = P;
  if (inverse) {
    B2_CUDA(cudaFuncSetAttribute(fluid_smem_kernel<H, W, NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fluid_smem_kernel<H, W, NT, true><<<(unsigned)grid, NT, smem, st>>>(f, out, (int)P, fp);
  } else {
    B2_CUDA(cudaFuncSetAttribute(fluid_smem_kernel<H, W, NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fluid_smem_kernel<H, W, NT, false><<<(unsigned)grid, NT, smem, st>>>(f, out, (int)P, fp);
  }
  B2_CHECK_LAUNCH();
  return B2_OK;
}

// ------------------------------------------------------------------ path B
#ifndef B2_FFT_ROWS
#define B2_FFT_ROWS 32
#endif
#ifndef B2_COLS_NT
#define B2_COLS_NT 256
#endif
constexpr int kRowsPerCta = B2_FFT_ROWS;
constexpr int kNTB = 256;
constexpr int kNTC = B2_COLS_NT;   // threads of the column kernel

template <int W, int DIR>
__global__ void __launch_bounds__(kNTB)
fft_rows_kernel(const float* __restrict__ f, float2* __restrict__ zg, float* __restrict__ out, int H) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int LD = W + 1, RB = kRowsPerCta;
  float2* z = reinterpret_cast<float2*>(smem_raw);
  float2* tw = z + RB * LD;
  const int tid = threadIdx.x, p = blockIdx.y, row0 = blockIdx.x * RB;
  const size_t N = (size_t)H * W;
  init_twiddles<W>(tw, tid, kNTB);
  float2* zp = zg + (size_t)p * N + (size_t)row0 * W;
  if (DIR < 0) {
    const float* f0 = f + (size_t)p * 2 * N + (size_t)row0 * W;
    const float* f1 = f0 + N;
    for (int i = tid; i < RB * W; i += kNTB) z[(i / W) * LD + (i % W)] = make_float2(f0[i], f1[i]);
  } else {
    for (int i = tid; i < RB * W; i += kNTB) z[(i / W) * LD + (i % W)] = zp[i];
  }
  __syncthreads();
  fft_lines<W, RB, DIR, kNTB, 1, LD>(z, tw, tid);
  if (DIR < 0) {
    for (int i = tid; i < RB * W; i += kNTB) zp[i] = z[(i / W) * LD + (i % W)];
  } else {
    float* o0 = out + (size_t)p * 2 * N + (size_t)row0 * W;
    float* o1 = o0 + N;
    for (int i = tid; i < RB * W; i += kNTB) {
      const float2 v = z[(i / W) * LD + (i % W)];
      o0[i] = v.x;
      o1[i] = v.y;
    }
  }
}

// Column FFT + multiply + column IFFT on column-cell block j and its mirror block.
template <int H, int W, bool INVERSE>
__global__ void __launch_bounds__(kNTC)
fft_cols_kernel(float2* __restrict__ zg, FluidParams fp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int N1w = Fact<W>::N1, N2w = Fact<W>::N2, NC = 2 * N2w, LD = NC + 1;
  float2* z = reinterpret_cast<float2*>(smem_raw);
  float2* tw = z + H * LD;
  float2* csH = tw + H;
  float2* csW = csH + H;
  const int tid = threadIdx.x, p = blockIdx.y, j = blockIdx.x, jm = (N1w - j) % N1w;
  const bool self = (j == jm);
  init_twiddles<H>(tw, tid, kNTC);
  init_symbol_lut<H>(csH, tid, kNTC);
  init_symbol_lut<W>(csW, tid, kNTC);
  float2* zp = zg + (size_t)p * H * W;
  for (int i = tid; i < H * NC; i += kNTC) {
    const int r = i / NC, cc = i % NC;
    const int col = (cc < N2w) ? j * N2w + cc : jm * N2w + (cc - N2w);
    z[r * LD + cc] = zp[(size_t)r * W + col];
  }
  __syncthreads();
  fft_lines<H, NC, -1, kNTC, LD, 1>(z, tw, tid);
  for (int t = tid; t < H * N2w; t += kNTC) {
    const int pr = t / N2w, cc = t % N2w, pc = j * N2w + cc;
    const int k0 = cell_to_freq<H>(pr), k1 = cell_to_freq<W>(pc);
    const int qr = freq_to_cell<H>((H - k0) & (H - 1)), qc = freq_to_cell<W>((W - k1) & (W - 1));
    const int qcc = qc - jm * N2w + (self ? 0 : N2w);
    const int lin = pr * W + pc, linq = qr * W + qc;
    if (self && lin > linq) continue;
    float A, Br, Bi;
    fluid_coeffs<INVERSE>(fp, csH[k0], csW[k1], A, Br, Bi);
    const float2 Z = z[pr * LD + cc];
    const float2 Zq = z[qr * LD + qcc];
    z[pr * LD + cc] = make_float2(A * Z.x + Br * Zq.x + Bi * Zq.y, A * Z.y + Bi * Zq.x - Br * Zq.y);
    if (lin != linq)
      z[qr * LD + qcc] = make_float2(A * Zq.x + Br * Z.x + Bi * Z.y, A * Zq.y + Bi * Z.x - Br * Z.y);
  }
  __syncthreads();
  fft_lines<H, NC, +1, kNTC, LD, 1>(z, tw, tid);
  const int ncols = self ? N2w : NC;
  for (int i = tid; i < H * NC; i += kNTC) {
    const int r = i / NC, cc = i % NC;
    if (cc >= ncols) continue;
    const int col = (cc < N2w) ? j * N2w + cc : jm * N2w + (cc - N2w);
    zp[(size_t)r * W + col] = z[r * LD + cc];
  }
}

template <int H, int W>
static int launch_3pass(const float* f, float* out, int64_t P, FluidParams fp, int inverse, float2* zg, cudaStream_t st) {
  constexpr int N1w = Fact<W>::N1, N2w = Fact<W>::N2;
  const size_t smem_rows = sizeof(float2) * ((size_t)kRowsPerCta * (W + 1) + W);
  const size_t smem_cols = sizeof(float2) * ((size_t)H * (2 * N2w + 1) + 2 * H + W);
  static_assert(H % kRowsPerCta == 0, "rows per CTA must divide H");
  for (int64_t p0 = 0; p0 < P; p0 += kMaxGridY) {
    const int64_t pn = (P - p0 < kMaxGridY) ? P - p0 : kMaxGridY;
    const float* fp0 = f + (size_t)p0 * 2 * H * W;
    float* op0 = out + (size_t)p0 * 2 * H * W;
    float2* zp0 = zg + (size_t)p0 * H * W;
    B2_CUDA(cudaFuncSetAttribute(fft_rows_kernel<W, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
    B2_CUDA(cudaFuncSetAttribute(fft_rows_kernel<W, +1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
    fft_rows_kernel<W, -1><<<dim3(H / kRowsPerCta, (unsigned)pn), kNTB, smem_rows, st>>>(fp0, zp0, nullptr, H);
    B2_CHECK_LAUNCH();
    if (inverse) {
      B2_CUDA(cudaFuncSetAttribute(fft_cols_kernel<H, W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
      fft_cols_kernel<H, W, true><<<dim3(N1w / 2 + 1, (unsigned)pn), kNTC, smem_cols, st>>>(zp0, fp);
    } else {
      B2_CUDA(cudaFuncSetAttribute(fft_cols_kernel<H, W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
      fft_cols_kernel<H, W, false><<<dim3(N1w / 2 + 1, (unsigned)pn), kNTC, smem_cols, st>>>(zp0, fp);
    }
    B2_CHECK_LAUNCH();
    fft_rows_kernel<W, +1><<<dim3(H / kRowsPerCta, (unsigned)pn), kNTB, smem_rows, st>>>(nullptr, zp0, op0, H);
    B2_CHECK_LAUNCH();
  }
  return B2_OK;
}

// ------------------------------------------------------------------ path B, fused EPDiff-step row kernels
// For grids that do not fit one SM (256x256, rectangular) an EPDiff step is three kernels instead of five:
//   adstar_rows_kernel : m = Ad*_u m0 computed per row band (u band + halo staged in shared memory, stencil
//                        from shared memory, m0 gathered from global) and row-FFT'd in place -> spectrum scratch
//   fft_cols_kernel    : column FFT + symbol multiply + column IFFT (above)
//   compose_rows_kernel: row IFFT -> v, then u_next = interp(u, v, -dt) - dt v gathered from global
// so m and v never round-trip through HBM (64*N instead of 96*N bytes per step).
#ifndef B2_BAND_ROWS
#define B2_BAND_ROWS 8
#endif
#ifndef B2_BAND_MINBLOCKS
#define B2_BAND_MINBLOCKS 6
#endif
constexpr int kBandRows = B2_BAND_ROWS;

template <int W, int BG>
__global__ void __launch_bounds__(kNTB, B2_BAND_MINBLOCKS)
adstar_rows_kernel(const float* __restrict__ u, const float* __restrict__ m0, float2* __restrict__ zg, int H) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int LD = W + 1, RB = kBandRows;
  float2* z = reinterpret_cast<float2*>(smem_raw);
  float2* tw = z + RB * LD;
  float2* ub = tw + W;                       // (RB+2) x W interleaved u rows, edge rows replicated
  const int tid = threadIdx.x, p = blockIdx.y, row0 = blockIdx.x * RB;
  const int N = H * W;
  init_twiddles<W>(tw, tid, kNTB);
  const float* u0 = u + (size_t)p * 2 * N;
  const float* u1 = u0 + N;
  for (int i = tid; i < (RB + 2) * W; i += kNTB) {
    const int k = i / W, c = i % W;
    const int r = min(max(row0 - 1 + k, 0), H - 1);
    ub[i] = make_float2(u0[r * W + c], u1[r * W + c]);
  }
  __syncthreads();
  const float* mp = m0 + (size_t)p * 2 * N;
  for (int i = tid; i < RB * W; i += kNTB) {
    const int rr = i / W, c = i % W, r = row0 + rr;
    const float2 ce = ub[(rr + 1) * W + c], up = ub[rr * W + c], dn = ub[(rr + 2) * W + c];
    const float2 lf = ub[(rr + 1) * W + max(c - 1, 0)], rt = ub[(rr + 1) * W + min(c + 1, W - 1)];
    const float sr = (r == 0 || r == H - 1) ? 1.f : 0.5f, sc = (c == 0 || c == W - 1) ? 1.f : 0.5f;
    const float d00 = sr * (dn.x - up.x), d10 = sr * (dn.y - up.y);
    const float d01 = sc * (rt.x - lf.x), d11 = sc * (rt.y - lf.y);
    const Taps t = make_taps_fwd<BG>((float)r + ce.x, (float)c + ce.y, H, W);
    const float w0 = tap_sample<BG>(t, mp[t.o00], mp[t.o10], mp[t.o01], mp[t.o11]);
    const float w1 = tap_sample<BG>(t, mp[N + t.o00], mp[N + t.o10], mp[N + t.o01], mp[N + t.o11]);
    z[rr * LD + c] = make_float2(w0 + (d00 * w0 + d10 * w1), w1 + (d01 * w0 + d11 * w1));
  }
  __syncthreads();
  fft_lines<W, RB, -1, kNTB, 1, LD>(z, tw, tid);
  float2* zp = zg + (size_t)p * N + (size_t)row0 * W;
  for (int i = tid; i < RB * W; i += kNTB) zp[i] = z[(i / W) * LD + (i % W)];
}

template <int W, int BG>
__global__ void __launch_bounds__(kNTB, B2_BAND_MINBLOCKS)
compose_rows_kernel(const float2* __restrict__ zg, const float* __restrict__ u, float* __restrict__ unext,
                    float* __restrict__ vout, int H, float mdt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int LD = W + 1, RB = kBandRows;
  float2* z = reinterpret_cast<float2*>(smem_raw);
  float2* tw = z + RB * LD;
  const int tid = threadIdx.x, p = blockIdx.y, row0 = blockIdx.x * RB;
  const int N = H * W;
  init_twiddles<W>(tw, tid, kNTB);
  const float2* zp = zg + (size_t)p * N + (size_t)row0 * W;
  for (int i = tid; i < RB * W; i += kNTB) z[(i / W) * LD + (i % W)] = zp[i];
  __syncthreads();
  fft_lines<W, RB, +1, kNTB, 1, LD>(z, tw, tid);
  const float* u0 = u ? u + (size_t)p * 2 * N : nullptr;
  float* n0 = unext + (size_t)p * 2 * N + (size_t)row0 * W;
  float* v0 = vout ? vout + (size_t)p * 2 * N + (size_t)row0 * W : nullptr;
  for (int i = tid; i < RB * W; i += kNTB) {
    const int rr = i / W, c = i % W, r = row0 + rr;
    const float2 v = z[rr * LD + c];
    float a = mdt * v.x, b = mdt * v.y;
    if (u0) {
      const Taps t = make_taps_fwd<BG>((float)r + a, (float)c + b, H, W);
      a += tap_sample<BG>(t, u0[t.o00], u0[t.o10], u0[t.o01], u0[t.o11]);
      b += tap_sample<BG>(t, u0[N + t.o00], u0[N + t.o10], u0[N + t.o01], u0[N + t.o11]);
    }
    n0[i] = a;
    n0[N + i] = b;
    if (v0) { v0[i] = v.x; v0[N + i] = v.y; }
  }
}

template <int H, int W>
static int big_step(const float* u, const float* m0, float* unext, float* vout, float2* zg, int64_t P,
                    FluidParams fp, float mdt, int bg, cudaStream_t st) {
  constexpr int N1w = Fact<W>::N1, N2w = Fact<W>::N2, RB = kBandRows;
  static_assert(H % RB == 0 && H % kRowsPerCta == 0, "row bands must divide H");
  const size_t smem_fft = sizeof(float2) * ((size_t)kRowsPerCta * (W + 1) + W);
  const size_t smem_ad = sizeof(float2) * ((size_t)RB * (W + 1) + W + (size_t)(RB + 2) * W);
  const size_t smem_co = sizeof(float2) * ((size_t)RB * (W + 1) + W);
  const size_t smem_cols = sizeof(float2) * ((size_t)H * (2 * N2w + 1) + 2 * H + W);
  for (int64_t p0 = 0; p0 < P; p0 += kMaxGridY) {
    const unsigned pn = (unsigned)((P - p0 < kMaxGridY) ? P - p0 : kMaxGridY);
    const size_t foff = (size_t)p0 * 2 * H * W;
    float2* zp0 = zg + (size_t)p0 * H * W;
    if (u) {
      if (bg == B2_BG_CLAMP) {
        B2_CUDA(cudaFuncSetAttribute(adstar_rows_kernel<W, B2_BG_CLAMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ad));
        adstar_rows_kernel<W, B2_BG_CLAMP><<<dim3(H / RB, pn), kNTB, smem_ad, st>>>(u + foff, m0 + foff, zp0, H);
      } else {
        B2_CUDA(cudaFuncSetAttribute(adstar_rows_kernel<W, B2_BG_ZERO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ad));
        adstar_rows_kernel<W, B2_BG_ZERO><<<dim3(H / RB, pn), kNTB, smem_ad, st>>>(u + foff, m0 + foff, zp0, H);
      }
    } else {   // u = 0: Ad* is the identity, m = m0
      B2_CUDA(cudaFuncSetAttribute(fft_rows_kernel<W, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fft));
      fft_rows_kernel<W, -1><<<dim3(H / kRowsPerCta, pn), kNTB, smem_fft, st>>>(m0 + foff, zp0, nullptr, H);
    }
    B2_CHECK_LAUNCH();
    B2_CUDA(cudaFuncSetAttribute(fft_cols_kernel<H, W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols));
    fft_cols_kernel<H, W, true><<<dim3(N1w / 2 + 1, pn), kNTC, smem_cols, st>>>(zp0, fp);
    B2_CHECK_LAUNCH();
    if (bg == B2_BG_CLAMP) {
      B2_CUDA(cudaFuncSetAttribute(compose_rows_kernel<W, B2_BG_CLAMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_co));
      compose_rows_kernel<W, B2_BG_CLAMP><<<dim3(H / RB, pn), kNTB, smem_co, st>>>(zp0, u ? u + foff : nullptr, unext + foff, vout ? vout + foff : nullptr, H, mdt);
    } else {
      B2_CUDA(cudaFuncSetAttribute(compose_rows_kernel<W, B2_BG_ZERO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_co));
      compose_rows_kernel<W, B2_BG_ZERO><<<dim3(H / RB, pn), kNTB, smem_co, st>>>(zp0, u ? u + foff : nullptr, unext + foff, vout ? vout + foff : nullptr, H, mdt);
    }
    B2_CHECK_LAUNCH();
  }
  return B2_OK;
}

static bool smem_path(int64_t H, int64_t W) { return H == W && (H == 16 || H == 32 || H == 64 || H == 128); }
static bool pass3_path(int64_t H, int64_t W) {
  return (H == 256 && W == 256) || (H == 64 && W == 128) || (H == 128 && W == 64) || (H == 256 && W == 128) ||
         (H == 128 && W == 256);
}

int fluid_apply_impl(const float* f, float* out, int64_t P, int64_t H, int64_t W, float alpha, float beta,
                     float gamma, int inverse, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  FluidParams fp{alpha, beta, gamma, 1.0f / (float)(H * W)};
  if (smem_path(H, W)) {
    switch ((int)H) {
      case 16: return launch_smem<16, 16, 128>(f, out, P, fp, inverse, st);
      case 32: return launch_smem<32, 32, 256>(f, out, P, fp, inverse, st);
      case 64: return launch_smem<64, 64, 256>(f, out, P, fp, inverse, st);
      case 128: return launch_smem<128, 128, 1024>(f, out, P, fp, inverse, st);
    }
  }
  if (pass3_path(H, W)) {
    if (!workspace || workspace_bytes < b2_fluid_workspace_bytes(P, H, W)) return B2_E_WORKSPACE;
    float2* zg = reinterpret_cast<float2*>(workspace);
    if (H == 256 && W == 256) return launch_3pass<256, 256>(f, out, P, fp, inverse, zg, st);
    if (H == 64 && W == 128) return launch_3pass<64, 128>(f, out, P, fp, inverse, zg, st);
    if (H == 128 && W == 64) return launch_3pass<128, 64>(f, out, P, fp, inverse, zg, st);
    if (H == 256 && W == 128) return launch_3pass<256, 128>(f, out, P, fp, inverse, zg, st);
    if (H == 128 && W == 256) return launch_3pass<128, 256>(f, out, P, fp, inverse, zg, st);
  }
  return B2_E_FFTSIZE;
}

// One EPDiff step on the big-grid path: unext = compose(u, sharp(Ad*_u m0), -dt); u == nullptr means u = 0.
// vout (optional) receives v = sharp(Ad*_u m0).  zg: P*H*W float2 scratch.
int epdiff_step_big(const float* u, const float* m0, float* unext, float* vout, void* zg, int64_t P, int64_t H,
                    int64_t W, float alpha, float beta, float gamma, float dt, int bg, cudaStream_t st) {
  FluidParams fp{alpha, beta, gamma, 1.0f / (float)(H * W)};
  float2* z = reinterpret_cast<float2*>(zg);
  if (H == 256 && W == 256) return big_step<256, 256>(u, m0, unext, vout, z, P, fp, -dt, bg, st);
  if (H == 64 && W == 128) return big_step<64, 128>(u, m0, unext, vout, z, P, fp, -dt, bg, st);
  if (H == 128 && W == 64) return big_step<128, 64>(u, m0, unext, vout, z, P, fp, -dt, bg, st);
  if (H == 256 && W == 128) return big_step<256, 128>(u, m0, unext, vout, z, P, fp, -dt, bg, st);
  if (H == 128 && W == 256) return big_step<128, 256>(u, m0, unext, vout, z, P, fp, -dt, bg, st);
  return B2_E_FFTSIZE;
}

}  // namespace b2

using namespace b2;

extern "C" int64_t b2_fluid_workspace_bytes(int64_t P, int64_t H, int64_t W) {
  if (P <= 0 || H <= 0 || W <= 0) return 0;
  if (smem_path(H, W)) return 0;
  return (int64_t)sizeof(float2) * P * H * W;
}

extern "C" int b2_fluid_apply(const float* f, float* out, int64_t P, int64_t H, int64_t W, float alpha, float beta,
                              float gamma, int inverse, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!f || !out) return B2_E_NULL;
  if (P <= 0 || P > ((int64_t)1 << 30)) return B2_E_SHAPE;
  if (!(gamma > 0.f) || alpha < 0.f || beta < 0.f) return B2_E_PARAM;
  if (!smem_path(H, W) && !pass3_path(H, W)) return B2_E_FFTSIZE;
  return fluid_apply_impl(f, out, P, H, W, alpha, beta, gamma, inverse, workspace, workspace_bytes, (cudaStream_t)stream);
}

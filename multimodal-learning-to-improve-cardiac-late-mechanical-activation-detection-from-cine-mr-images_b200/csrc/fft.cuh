// Batched shared-memory FFT + Fourier-domain fluid-metric multiplier (sm_100a).
// Replaces torch.rfft/irfft (cuFFT) + lagomorph_ext.fluid_operator.
//
// The two real components of a vector field are packed as ONE complex field
// z = f0 + i*f1, so a single complex 2-D FFT carries both spectra:
//   F0(k) = (Z(k) + conj Z(-k))/2,  F1(k) = (Z(k) - conj Z(-k))/(2i).
// Applying the real symmetric 2x2 symbol [[a,b],[b,d]](k) (even in k) and
// re-packing gives   W(k) = A Z(k) + B conj Z(-k),  A = (a+d)/2, B = (a-d)/2 + i b,
// so the multiplier couples storage cell k with its mirror -k and nothing else.
//
// Each 1-D FFT of length N = N1*N2 runs as two in-register DFT passes (radix
// N1 then N2, both in {4,8,16}) that are IN PLACE per thread; the forward
// transform leaves the spectrum in a digit-permuted order (cell N2*k1+k2 holds
// frequency k1+N1*k2) which the multiplier decodes and the inverse consumes, so
// no reordering pass exists.  One routine serves rows and columns: lanes run
// across "lines" and each thread's points are strided; with the odd row pitch
// W+1 (in float2) both directions are shared-memory bank-conflict free.
#pragma once
#include "common.cuh"

namespace b2 {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// x * (c + i*s)
__device__ __forceinline__ float2 cmul(float2 x, float c, float s) {
  return make_float2(x.x * c - x.y * s, x.x * s + x.y * c);
}

// x * exp(DIR * 2*pi*i * t/16), t compile-time after unrolling
template <int DIR>
__device__ __forceinline__ float2 tw16(float2 x, int t) {
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R2 = 0.70710678118654752f;
  switch (t & 15) {
    case 0: return x;
    case 4: return DIR > 0 ? make_float2(-x.y, x.x) : make_float2(x.y, -x.x);
    case 8: return make_float2(-x.x, -x.y);
    case 12: return DIR > 0 ? make_float2(x.y, -x.x) : make_float2(-x.y, x.x);
    case 1: return cmul(x, C1, DIR * S1);
    case 2: return cmul(x, R2, DIR * R2);
    case 3: return cmul(x, S1, DIR * C1);
    case 5: return cmul(x, -S1, DIR * C1);
    case 6: return cmul(x, -R2, DIR * R2);
    case 7: return cmul(x, -C1, DIR * S1);
    default: return cmul(x, 1.f, 0.f);  // t in 9..15 never generated (k < R/2)
  }
}

#ifndef B2_FFT_FMA_BUTTERFLY
#define B2_FFT_FMA_BUTTERFLY 1
#endif
// Radix-2 butterfly (x0, x1) = (e + w o, e - w o) with w = exp(DIR 2 pi i t / 16), t compile-time after unrolling.
// For the non-trivial twiddles the product is never formed: x0 = e + w o as two chained FMAs per component and
// x1 = 2 e - x0 as one - 6 instructions instead of 4 (complex multiply) + 4 (add, subtract).
template <int DIR>
__device__ __forceinline__ void bfly16(float2 e, float2 o, int t, float2& x0, float2& x1) {
#if B2_FFT_FMA_BUTTERFLY
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R2 = 0.70710678118654752f;
  float c = 1.f, s = 0.f;
  switch (t & 15) {
    case 1: c = C1; s = DIR * S1; break;
    case 2: c = R2; s = DIR * R2; break;
    case 3: c = S1; s = DIR * C1; break;
    case 5: c = -S1; s = DIR * C1; break;
    case 6: c = -R2; s = DIR * R2; break;
    case 7: c = -C1; s = DIR * S1; break;
    default: {
      const float2 w = tw16<DIR>(o, t);
      x0 = cadd(e, w);
      x1 = csub(e, w);
      return;
    }
  }
  x0.x = fmaf(o.x, c, fmaf(-o.y, s, e.x));
  x0.y = fmaf(o.x, s, fmaf(o.y, c, e.y));
  x1.x = fmaf(2.f, e.x, -x0.x);
  x1.y = fmaf(2.f, e.y, -x0.y);
#else
  const float2 w = tw16<DIR>(o, t);
  x0 = cadd(e, w);
  x1 = csub(e, w);
#endif
}

// In-register DFT of R points, natural order in and out (radix-2 DIT recursion,
// fully unrolled; twiddles are compile-time constants).
template <int R, int DIR>
struct DftReg {
  static __device__ __forceinline__ void run(float2 (&x)[R]) {
    float2 e[R / 2], o[R / 2];
#pragma unroll
    for (int k = 0; k < R / 2; ++k) { e[k] = x[2 * k]; o[k] = x[2 * k + 1]; }
    DftReg<R / 2, DIR>::run(e);
    DftReg<R / 2, DIR>::run(o);
#pragma unroll
    for (int k = 0; k < R / 2; ++k) bfly16<DIR>(e[k], o[k], k * (16 / R), x[k], x[k + R / 2]);
  }
};
template <int DIR>
struct DftReg<1, DIR> {
  static __device__ __forceinline__ void run(float2 (&)[1]) {}
};

template <int N> struct Fact;
template <> struct Fact<16>  { static constexpr int N1 = 4,  N2 = 4;  };
template <> struct Fact<32>  { static constexpr int N1 = 8,  N2 = 4;  };
template <> struct Fact<64>  { static constexpr int N1 = 8,  N2 = 8;  };
template <> struct Fact<128> { static constexpr int N1 = 16, N2 = 8;  };
template <> struct Fact<256> { static constexpr int N1 = 16, N2 = 16; };

// storage cell <-> frequency of the digit-permuted spectrum
template <int N> __device__ __forceinline__ int cell_to_freq(int p) {
  return (p / Fact<N>::N2) + Fact<N>::N1 * (p % Fact<N>::N2);
}
template <int N> __device__ __forceinline__ int freq_to_cell(int k) {
  return Fact<N>::N2 * (k % Fact<N>::N1) + (k / Fact<N>::N1);
}

// tw[j] = (cos(2 pi j/N), sin(2 pi j/N)), j in [0, N)
template <int N>
__device__ __forceinline__ void init_twiddles(float2* tw, int tid, int nthreads) {
  for (int j = tid; j < N; j += nthreads) {
    float s, c;
    sincospif(2.0f * (float)j / (float)N, &s, &c);
    tw[j] = make_float2(c, s);
  }
}

#ifndef B2_FFT_GROUPBAR
#define B2_FFT_GROUPBAR 0
#endif
// Barrier between the two radix passes of one 1-D transform.  A line is touched only by the threads with the same
// (tid % NL), in both passes (NT is a multiple of NL): the warps that share a block of 32 lines form a closed group, so a
// named barrier over that group (bar.sync id, count) is enough - the groups drift independently instead of all
// NT threads meeting.  Falls back to __syncthreads() when the lines do not split into whole warps.
template <int NL, int NT>
__device__ __forceinline__ void fft_pass_barrier(int tid) {
#if B2_FFT_GROUPBAR
  if constexpr (NL % 32 == 0 && NT % NL == 0 && (NL / 32) > 1 && (NL / 32) <= 15) {
    constexpr int groups = NL / 32, per = NT / groups;
    asm volatile("bar.sync %0, %1;" ::"r"(1 + (tid % NL) / 32), "r"(per) : "memory");
    return;
  }
#endif
  (void)tid;
  __syncthreads();
}

// FFTs of length N along element stride `es`, for NL lines spaced `ls` apart.
// Forward (DIR=-1): natural in -> permuted out.  Inverse (DIR=+1): permuted in ->
// natural out, unnormalised.  Ends with __syncthreads().
template <int N, int NL, int DIR, int NT, int ES, int LS>
__device__ __forceinline__ void fft_lines(float2* __restrict__ z, const float2* __restrict__ tw, int tid) {
  constexpr int N1 = Fact<N>::N1, N2 = Fact<N>::N2;
  constexpr int es = ES, ls = LS;   // compile-time strides: every smem access gets an immediate offset
  if (DIR < 0) {
    for (int t = tid; t < NL * N2; t += NT) {
      const int line = t % NL, n2 = t / NL;
      float2* base = z + line * ls + n2 * es;
      float2 x[N1];
#pragma unroll
      for (int n1 = 0; n1 < N1; ++n1) x[n1] = base[(N2 * n1) * es];
      DftReg<N1, -1>::run(x);
#pragma unroll
      for (int k1 = 1; k1 < N1; ++k1) {
        const float2 w = tw[n2 * k1];
        x[k1] = cmul(x[k1], w.x, -w.y);
      }
#pragma unroll
      for (int k1 = 0; k1 < N1; ++k1) base[(N2 * k1) * es] = x[k1];
    }
    fft_pass_barrier<NL, NT>(tid);
    for (int t = tid; t < NL * N1; t += NT) {
      const int line = t % NL, k1 = t / NL;
      float2* base = z + line * ls + (N2 * k1) * es;
      float2 y[N2];
#pragma unroll
      for (int n2 = 0; n2 < N2; ++n2) y[n2] = base[n2 * es];
      DftReg<N2, -1>::run(y);
#pragma unroll
      for (int k2 = 0; k2 < N2; ++k2) base[k2 * es] = y[k2];
    }
    __syncthreads();
  } else {
    for (int t = tid; t < NL * N1; t += NT) {
      const int line = t % NL, k1 = t / NL;
      float2* base = z + line * ls + (N2 * k1) * es;
      float2 y[N2];
#pragma unroll
      for (int k2 = 0; k2 < N2; ++k2) y[k2] = base[k2 * es];
      DftReg<N2, +1>::run(y);
#pragma unroll
      for (int n2 = 0; n2 < N2; ++n2) {
        const float2 w = tw[n2 * k1];
        base[n2 * es] = (n2 == 0) ? y[n2] : cmul(y[n2], w.x, w.y);
      }
    }
    fft_pass_barrier<NL, NT>(tid);
    for (int t = tid; t < NL * N2; t += NT) {
      const int line = t % NL, n2 = t / NL;
      float2* base = z + line * ls + n2 * es;
      float2 x[N1];
#pragma unroll
      for (int k1 = 0; k1 < N1; ++k1) x[k1] = base[(N2 * k1) * es];
      DftReg<N1, +1>::run(x);
#pragma unroll
      for (int n1 = 0; n1 < N1; ++n1) base[(N2 * n1) * es] = x[n1];
    }
    __syncthreads();
  }
}

// Per-axis symbol LUTs: cs[k] = (4 sin^2(pi k/N), sin(2 pi k/N)) = (2(1-cos), sin)
template <int N>
__device__ __forceinline__ void init_symbol_lut(float2* cs, int tid, int nthreads) {
  for (int k = tid; k < N; k += nthreads) {
    float s, c, sh, ch;
    sincospif(2.0f * (float)k / (float)N, &s, &c);
    sincospif((float)k / (float)N, &sh, &ch);
    cs[k] = make_float2(4.0f * sh * sh, s);
  }
}

struct FluidParams { float alpha, beta, gamma, scale; };

#ifndef B2_MULT_UNROLL
#define B2_MULT_UNROLL 2
#endif
constexpr int kMultUnroll = B2_MULT_UNROLL;

// A, B of W = A Z + B conj(Z~) at frequency (k0, k1)
template <bool INVERSE>
__device__ __forceinline__ void fluid_coeffs(const FluidParams& fp, float2 cs0, float2 cs1,
                                             float& A, float& Br, float& Bi) {
  const float lam = fp.gamma + fp.alpha * (cs0.x + cs1.x);
  const float L00 = lam + fp.beta * cs0.x;
  const float L11 = lam + fp.beta * cs1.x;
  const float L01 = fp.beta * (cs0.y * cs1.y);
  float a, d, b;
  if (INVERSE) {
    const float idet = 1.0f / (L00 * L11 - L01 * L01);
    a = L11 * idet; d = L00 * idet; b = -L01 * idet;
  } else {
    a = L00; d = L11; b = L01;
  }
  A = fp.scale * 0.5f * (a + d);
  Br = fp.scale * 0.5f * (a - d);
  Bi = fp.scale * b;
}

// In-place multiplier on the permuted 2-D spectrum z[r*(W+1) + c].  Work is enumerated over the
// canonical half of the row frequencies k0 in [0, H/2]: a thread owns cell (k0, pc) AND its mirror
// (-k0, -k1) and writes both, so nobody idles on a "not my pair" test.  The thread's column (pc), its
// mirror column and the column part of the symbol are loop invariant (NT is a multiple of W); the row
// part is warp uniform.  Ends with __syncthreads().
template <int H, int W, bool INVERSE, int NT>
__device__ __forceinline__ void fluid_multiply(float2* __restrict__ z, const float2* __restrict__ csH,
                                               const float2* __restrict__ csW, const FluidParams fp, int tid) {
  static_assert(NT % W == 0, "threads per CTA must be a multiple of the row length");
  constexpr int LD = W + 1, RB = NT / W;
  const int pc = tid % W, br = tid / W;
  const int k1 = cell_to_freq<W>(pc);
  const int qc = freq_to_cell<W>((W - k1) & (W - 1));
  const float2 cs1 = csW[k1];
  const float h = 0.5f * fp.scale;
  const float bs1 = fp.beta * cs1.y;
#pragma unroll (kMultUnroll)
  for (int k0 = br; k0 <= H / 2; k0 += RB) {
    const int pr = freq_to_cell<H>(k0), qr = freq_to_cell<H>((H - k0) & (H - 1));
    if (pr == qr && pc > qc) continue;          // self-mirrored rows (k0 = 0, H/2): each pair once
    const float2 cs0 = csH[k0];
    const float lam = fp.gamma + fp.alpha * (cs0.x + cs1.x);
    const float L00 = lam + fp.beta * cs0.x;
    const float L11 = lam + fp.beta * cs1.x;
    const float L01 = cs0.y * bs1;
    float A, Br, Bi;
    if (INVERSE) {
      const float idet = __fdividef(h, L00 * L11 - L01 * L01);
      A = idet * (L11 + L00); Br = idet * (L11 - L00); Bi = -2.0f * idet * L01;
    } else {
      A = h * (L00 + L11); Br = h * (L00 - L11); Bi = 2.0f * h * L01;
    }
    float2* zp = z + pr * LD + pc;
    float2* zq = z + qr * LD + qc;
    const float2 Z = *zp, Zq = *zq;
    // W(k) = A Z + B conj(Zq);  W(-k) = A Zq + B conj(Z)
    *zp = make_float2(A * Z.x + Br * Zq.x + Bi * Zq.y, A * Z.y + Bi * Zq.x - Br * Zq.y);
    if (zp != zq) *zq = make_float2(A * Zq.x + Br * Z.x + Bi * Z.y, A * Zq.y + Bi * Z.x - Br * Z.y);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// Fused middle of the operator: [second radix pass of the forward column FFT] + [symbol multiply] + [first radix pass
// of the inverse column FFT], all in registers.  The second forward pass of task (column pc, k1) leaves the N2
// frequencies k0 = k1 + N1 k2 of that column in the thread's registers, and the first inverse pass consumes exactly
// the same cells, so only the multiplier's partner - frequency (-k0, -k1f), i.e. task (mirror column qc,
// k1' = (N1 - k1) % N1) with register index k2' - has to be in the same thread.  Units of work are therefore PAIRS of
// tasks closed under mirroring; their number is (N1/2) W, one per thread for every instantiation used (128 x 128 with
// 1024 threads, 64 x 64 with 256).  Saves two shared-memory round trips of the whole field, two barriers and the
// multiplier's cell-index arithmetic per operator.
//   unit  <  (N1/2 - 1) W          : k1 in 1 .. N1/2-1, any column pc        <->  (qc, N1 - k1)
//   next  2 (W/2 - 1) units        : k1 in {0, N1/2}, column pairs pc < qc   <->  (qc, k1)
//   last  2 units                  : the two self-mirrored columns (k1f = 0, W/2) for k1 = 0 and k1 = N1/2
// ---------------------------------------------------------------------------------------------------------------
#ifndef B2_FUSED_MULT
#define B2_FUSED_MULT 1
#endif

template <int W>
__device__ __forceinline__ int mirror_cell(int pc) { return freq_to_cell<W>((W - cell_to_freq<W>(pc)) & (W - 1)); }

// canonical representative jj in [0, W/2 - 1) of the column pairs pc < qc (contiguous cells wherever possible)
template <int W>
__device__ __forceinline__ int canonical_column(int jj) {
  constexpr int N1 = Fact<W>::N1, N2 = Fact<W>::N2;
  constexpr int nA = (N1 / 2 - 1) * N2;              // whole blocks a = 1 .. N1/2-1
  if (jj < nA) return N2 + jj;
  const int j2 = jj - nA;
  if (j2 < N2 / 2 - 1) return 1 + j2;                // block 0: b = 1 .. N2/2-1  (b' = N2 - b)
  return N2 * (N1 / 2) + (j2 - (N2 / 2 - 1));        // block N1/2: b = 0 .. N2/2-1  (b' = N2 - 1 - b)
}

// W(k) = A Z(k) + B conj Z(-k) for one mirror pair held in registers
template <bool INVERSE>
__device__ __forceinline__ void mult_pair(float2& Z, float2& Zq, bool same, const FluidParams& fp, float2 cs0, float2 cs1) {
  const float h = 0.5f * fp.scale;
  const float lam = fp.gamma + fp.alpha * (cs0.x + cs1.x);
  const float L00 = lam + fp.beta * cs0.x, L11 = lam + fp.beta * cs1.x, L01 = fp.beta * (cs0.y * cs1.y);
  float A, Br, Bi;
  if (INVERSE) {
    const float idet = __fdividef(h, L00 * L11 - L01 * L01);
    A = idet * (L11 + L00); Br = idet * (L11 - L00); Bi = -2.0f * idet * L01;
  } else {
    A = h * (L00 + L11); Br = h * (L00 - L11); Bi = 2.0f * h * L01;
  }
  const float2 z = Z, zq = Zq;
  Z = make_float2(A * z.x + Br * zq.x + Bi * zq.y, A * z.y + Bi * zq.x - Br * zq.y);
  if (!same) Zq = make_float2(A * zq.x + Br * z.x + Bi * z.y, A * zq.y + Bi * z.x - Br * z.y);
}

// one task's share of the inverse column FFT's first pass: DFT over k2, inter-pass twiddle, store
template <int N1, int N2, int ES>
__device__ __forceinline__ void inv_pass1_store(float2 (&y)[N2], float2* base, const float2* tw, int k1) {
  DftReg<N2, +1>::run(y);
#pragma unroll
  for (int n2 = 0; n2 < N2; ++n2) {
    const float2 w = tw[n2 * k1];
    base[n2 * ES] = (n2 == 0) ? y[n2] : cmul(y[n2], w.x, w.y);
  }
}

template <int H, int W, bool INVERSE, int NT>
__device__ __forceinline__ void fluid_cols_mid_fused(float2* __restrict__ z, const float2* __restrict__ twH,
                                                     const float2* __restrict__ csH, const float2* __restrict__ csW,
                                                     const FluidParams fp, int tid) {
  constexpr int N1 = Fact<H>::N1, N2 = Fact<H>::N2, LD = W + 1, ES = LD;
  constexpr int nA = (N1 / 2 - 1) * W, nB = 2 * (W / 2 - 1), units = nA + nB + 2;
  static_assert(units == (N1 / 2) * W, "unit count");
  for (int unit = tid; unit < units; unit += NT) {
    if (unit < nA + nB) {
      // ---- a pair of distinct tasks (pc, k1) <-> (qc, k1q)
      int k1, pc;
      if (unit < nA) { k1 = 1 + unit / W; pc = unit % W; }
      else {
        const int j = unit - nA;
        k1 = (j < W / 2 - 1) ? 0 : N1 / 2;
        pc = canonical_column<W>(j < W / 2 - 1 ? j : j - (W / 2 - 1));
      }
      const int k1q = (N1 - k1) & (N1 - 1), qc = mirror_cell<W>(pc);
      float2* basea = z + pc + (N2 * k1) * ES;
      float2* baseb = z + qc + (N2 * k1q) * ES;
      float2 ya[N2], yb[N2];
#pragma unroll
      for (int n2 = 0; n2 < N2; ++n2) { ya[n2] = basea[n2 * ES]; yb[n2] = baseb[n2 * ES]; }
      DftReg<N2, -1>::run(ya);
      DftReg<N2, -1>::run(yb);
      const float2 cs1 = csW[cell_to_freq<W>(pc)];
      if (k1 == 0) {
#pragma unroll
        for (int k2 = 0; k2 < N2; ++k2) mult_pair<INVERSE>(ya[k2], yb[(N2 - k2) % N2], false, fp, csH[N1 * k2], cs1);
      } else {
#pragma unroll
        for (int k2 = 0; k2 < N2; ++k2) mult_pair<INVERSE>(ya[k2], yb[N2 - 1 - k2], false, fp, csH[k1 + N1 * k2], cs1);
      }
      inv_pass1_store<N1, N2, ES>(ya, basea, twH, k1);
      inv_pass1_store<N1, N2, ES>(yb, baseb, twH, k1q);
    } else {
      // ---- the self-mirrored columns (k1f = 0 and W/2): mirror partners inside one task's registers
      const int k1 = (unit == nA + nB) ? 0 : N1 / 2;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const int k1f = which ? W / 2 : 0, pc = freq_to_cell<W>(k1f);
        float2* base = z + pc + (N2 * k1) * ES;
        float2 y[N2];
#pragma unroll
        for (int n2 = 0; n2 < N2; ++n2) y[n2] = base[n2 * ES];
        DftReg<N2, -1>::run(y);
        const float2 cs1 = csW[k1f];
        if (k1 == 0) {
#pragma unroll
          for (int k2 = 0; k2 <= N2 / 2; ++k2)
            mult_pair<INVERSE>(y[k2], y[(N2 - k2) % N2], k2 == (N2 - k2) % N2, fp, csH[N1 * k2], cs1);
        } else {
#pragma unroll
          for (int k2 = 0; k2 < N2 / 2; ++k2) mult_pair<INVERSE>(y[k2], y[N2 - 1 - k2], false, fp, csH[k1 + N1 * k2], cs1);
        }
        inv_pass1_store<N1, N2, ES>(y, base, twH, k1);
      }
    }
  }
  __syncthreads();
}

// first pass of a forward transform / second pass of an inverse transform on their own (the halves of fft_lines
// that stay in front of and behind the fused middle)
template <int N, int NL, int NT, int ES, int LS>
__device__ __forceinline__ void fft_fwd_pass1(float2* __restrict__ z, const float2* __restrict__ tw, int tid) {
  constexpr int N1 = Fact<N>::N1, N2 = Fact<N>::N2;
  for (int t = tid; t < NL * N2; t += NT) {
    const int line = t % NL, n2 = t / NL;
    float2* base = z + line * LS + n2 * ES;
    float2 x[N1];
#pragma unroll
    for (int n1 = 0; n1 < N1; ++n1) x[n1] = base[(N2 * n1) * ES];
    DftReg<N1, -1>::run(x);
#pragma unroll
    for (int k1 = 1; k1 < N1; ++k1) {
      const float2 w = tw[n2 * k1];
      x[k1] = cmul(x[k1], w.x, -w.y);
    }
#pragma unroll
    for (int k1 = 0; k1 < N1; ++k1) base[(N2 * k1) * ES] = x[k1];
  }
  __syncthreads();
}
// second pass of a forward transform on its own (the cluster kernels run the first one on operands pulled over DSMEM)
template <int N, int NL, int NT, int ES, int LS>
__device__ __forceinline__ void fft_fwd_pass2(float2* __restrict__ z, int tid) {
  constexpr int N1 = Fact<N>::N1, N2 = Fact<N>::N2;
  for (int t = tid; t < NL * N1; t += NT) {
    const int line = t % NL, k1 = t / NL;
    float2* base = z + line * LS + (N2 * k1) * ES;
    float2 y[N2];
#pragma unroll
    for (int n2 = 0; n2 < N2; ++n2) y[n2] = base[n2 * ES];
    DftReg<N2, -1>::run(y);
#pragma unroll
    for (int k2 = 0; k2 < N2; ++k2) base[k2 * ES] = y[k2];
  }
  __syncthreads();
}
template <int N, int NL, int NT, int ES, int LS>
__device__ __forceinline__ void fft_inv_pass2(float2* __restrict__ z, int tid) {
  constexpr int N1 = Fact<N>::N1, N2 = Fact<N>::N2;
  for (int t = tid; t < NL * N2; t += NT) {
    const int line = t % NL, n2 = t / NL;
    float2* base = z + line * LS + n2 * ES;
    float2 x[N1];
#pragma unroll
    for (int k1 = 0; k1 < N1; ++k1) x[k1] = base[(N2 * k1) * ES];
    DftReg<N1, +1>::run(x);
#pragma unroll
    for (int n1 = 0; n1 < N1; ++n1) base[(N2 * n1) * ES] = x[n1];
  }
  __syncthreads();
}

// Whole operator on a field resident in shared memory as z = f0 + i f1.
template <int H, int W, bool INVERSE, int NT>
__device__ __forceinline__ void fluid_smem(float2* z, const float2* twH, const float2* twW,
                                           const float2* csH, const float2* csW, const FluidParams fp, int tid) {
  constexpr int LD = W + 1;
  fft_lines<W, H, -1, NT, 1, LD>(z, twW, tid);      // rows: FFT along c, lanes along r
#if B2_FUSED_MULT
  fft_fwd_pass1<H, W, NT, LD, 1>(z, twH, tid);      // cols: first radix pass
  fluid_cols_mid_fused<H, W, INVERSE, NT>(z, twH, csH, csW, fp, tid);
  fft_inv_pass2<H, W, NT, LD, 1>(z, tid);
#else
  fft_lines<H, W, -1, NT, LD, 1>(z, twH, tid);      // cols: FFT along r, lanes along c
  fluid_multiply<H, W, INVERSE, NT>(z, csH, csW, fp, tid);
  fft_lines<H, W, +1, NT, LD, 1>(z, twH, tid);
#endif
  fft_lines<W, H, +1, NT, 1, LD>(z, twW, tid);
}

template <int H, int W>
struct FluidSmem {
  static constexpr int LD = W + 1;
  static constexpr size_t z_bytes = sizeof(float2) * (size_t)H * LD;
  static constexpr size_t lut_bytes = sizeof(float2) * 2 * (size_t)(H + W);
  static constexpr size_t bytes = z_bytes + lut_bytes;
  // layout: z | twH[H] | twW[W] | csH[H] | csW[W]
  static __device__ __forceinline__ void carve(unsigned char* smem, float2*& z, float2*& twH, float2*& twW,
                                               float2*& csH, float2*& csW) {
    z = reinterpret_cast<float2*>(smem);
    twH = z + (size_t)H * LD;
    twW = twH + H;
    csH = twW + W;
    csW = csH + H;
  }
  static __device__ __forceinline__ void init_luts(float2* twH, float2* twW, float2* csH, float2* csW,
                                                   int tid, int nthreads) {
    init_twiddles<H>(twH, tid, nthreads);
    init_twiddles<W>(twW, tid, nthreads);
    init_symbol_lut<H>(csH, tid, nthreads);
    init_symbol_lut<W>(csW, tid, nthreads);
  }
};

}  // namespace b2

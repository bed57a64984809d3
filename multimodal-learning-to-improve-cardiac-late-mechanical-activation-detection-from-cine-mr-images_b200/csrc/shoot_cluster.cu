// Fused geodesic shooting for 256x256 grids: one 4-CTA thread-block cluster per frame-pair.
//
// A 256x256 complex field is 512 KiB - it does not fit one SM - so the pair is split into four 64-row slabs, one per
// CTA of a cluster (4 SMs).  Row FFTs run on the slab in shared memory; for the column pass the four CTAs exchange
// data through a per-pair scratch that stays in L2: every CTA stores its slab, the cluster synchronises, and every
// CTA loads a MIRROR-CLOSED group of 64 spectrum columns (4 blocks of 16 cells: {0,8,1,15}, {2,14,3,13}, ...), so
// the k <-> -k coupling of the fluid multiplier stays inside one CTA.  m / v never reach HBM; u_s and m0 live in a
// per-cluster L2 scratch and are re-read by the gathers; the exchange buffer is the field u_{s+1} is about to be
// written into, so a cluster keeps 1.5 MiB live (50 MiB for the 33 resident clusters).
// Three cluster barriers per EPDiff step (after compose, after the row pass, after the column pass).
#include <cooperative_groups.h>

#include "fft.cuh"
#include "shoot_params.cuh"
#include "strain.cuh"

namespace cg = cooperative_groups;

namespace b2 {

// Cluster shape (compile-time): 4 CTAs x 1024 threads with 64-row slabs (one CTA per SM), or 8 CTAs x 512 threads with
// 32-row slabs (two CTAs per SM: two barrier domains per SM, and 8-CTA clusters of half-size CTAs pack the GPCs'
// SM counts better than 4 whole SMs do).  Work per thread is identical: 16 pixels, one radix-16 task per FFT pass.
#ifndef B2_CLUSTER_CTAS
#define B2_CLUSTER_CTAS 4
#endif
constexpr int kCH = 256, kCW = 256, kCL = B2_CLUSTER_CTAS, kCNT = 4096 / kCL, kSR = kCH / kCL;   // slab rows per CTA
constexpr int kCtasPerSM = kCL / 4;                          // launch bound: 1 (1024 threads) or 2 (512 threads)
static_assert(kCL == 4 || kCL == 8, "cluster of 4 or 8 CTAs");
constexpr int kLDR = kCW + 1;      // row-slab pitch   (64 rows x 257)
constexpr int kLDC = kSR + 1;      // column-slab pitch (256 rows x 65)
constexpr int kCN = kCH * kCW;

// mirror-closed groups of spectrum-cell blocks (16 cells each; the mirror of block b is block (16 - b) % 16)
#if B2_CLUSTER_CTAS == 4
__constant__ int c_group_block[4][4] = {{0, 8, 1, 15}, {2, 14, 3, 13}, {4, 12, 5, 11}, {6, 10, 7, 9}};
__device__ __forceinline__ int mirror_q(int g, int q) { return g == 0 ? (q < 2 ? q : 5 - q) : (q ^ 1); }
#else
__constant__ int c_group_block[8][2] = {{0, 8}, {1, 15}, {2, 14}, {3, 13}, {4, 12}, {5, 11}, {6, 10}, {7, 9}};
__device__ __forceinline__ int mirror_q(int g, int q) { return g == 0 ? q : (q ^ 1); }     // blocks 0 and 8 mirror onto themselves
#endif

struct ClusterParams {
  b2_shoot_args a;
  float* scratch;        // per cluster: [u ping | u pong | m0] + bins; the spectrum exchange reuses the next-u field
  int64_t P;                // pairs of the whole batch (trajectory strides)
  int64_t p0, p1;           // pairs this launch processes: [p0, p1)
  int64_t cluster_stride;   // floats of scratch per cluster
};

// The spectrum exchange buffer Zs (one complex field) has no memory of its own: it lives in the 2-plane field that
// the coming compose will overwrite with u_{s+1} (dead until then).  Row r of the spectrum is placed so that the rows
// CTA X reads LAST (its own slab, second exchange) occupy exactly the bytes X itself writes in that compose - rows
// 64X .. 64X+31 in X's part of plane 0, rows 64X+32 .. 64X+63 in X's part of plane 1 - so no CTA can overwrite
// spectrum rows another CTA still has to read, and the cluster needs no extra barrier.  Index in float2 units.
__device__ __forceinline__ int zs_row(int r) {
  constexpr int half = kSR / 2;
  const int x = (int)((unsigned)r / (unsigned)kSR), lr = (int)((unsigned)r % (unsigned)kSR);     // kSR is a power of two
  return (lr < half ? 0 : kCN / 2 - half * kCW) + x * (half * kCW) + lr * kCW;
}

#ifndef B2_CLUSTER_DSMEM
#define B2_CLUSTER_DSMEM 0     // 1: row <-> column exchange over distributed shared memory (measured: 16.55 vs 11.87 ms)
#endif
// ---- distributed shared memory: the slab of CTA g of this cluster, read in place
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned dsmem_of_rank(unsigned addr, int rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float2 ld_dsmem2(unsigned addr) {
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
// home of spectrum-cell block b (16 cells): CTA g, local block q  <->  c_group_block[g][q] == b
__device__ __forceinline__ void block_home(int b, int& g, int& q) {
  g = 0; q = 0;
#pragma unroll
  for (int gg = 0; gg < kCL; ++gg)
#pragma unroll
    for (int qq = 0; qq < 16 / kCL; ++qq)
      if (c_group_block[gg][qq] == b) { g = gg; q = qq; }
}

// One fluid operator on the slab in z (row layout in, row layout out): row FFT -> spectrum exchange through the
// per-cluster L2 scratch Zs -> column FFT of a mirror-closed column group -> multiplier -> inverse, same way back.
// A plain inlined function (a by-reference lambda put its closure in local memory).
#if !B2_CLUSTER_DSMEM
template <bool inverse>
__device__ __forceinline__ void cluster_fluid(cg::cluster_group& cluster, float2* __restrict__ z, const float2* tw,
                                              const float2* cs, float2* Zs, const FluidParams fp, const int tid,
                                              const int rk, const int r0, const int c, const int br, const int lc,
                                              const int q, const int pc) {
  constexpr int H = kCH, W = kCW, RBc = kCNT / kCW, NBc = kSR / RBc;
  fft_lines<256, kSR, -1, kCNT, 1, kLDR>(z, tw, tid);
  for (int k = 0; k < NBc; ++k) {                   // slab -> L2 scratch (cell order along c)
    const int lr = k * RBc + br;
    Zs[zs_row(r0 + lr) + c] = z[lr * kLDR + c];
  }
  cluster.sync();   // release/acquire at cluster scope: global writes of the other CTAs are visible
  for (int i = tid; i < H * kSR; i += kCNT) {       // mirror-closed column group, all 256 rows
    const int r = i / kSR;
    z[r * kLDC + lc] = Zs[zs_row(r) + pc];
  }
  __syncthreads();
  fft_lines<256, kSR, -1, kCNT, kLDC, 1>(z, tw, tid);
  {
    const int k1 = cell_to_freq<256>(pc);
    const int qcell = freq_to_cell<256>((W - k1) & (W - 1));
    const int lc2 = mirror_q(rk, q) * 16 + (qcell & 15);
    const float2 cs1 = cs[k1];
    const float h = 0.5f * fp.scale, bs1 = fp.beta * cs1.y;
    for (int k0 = tid / kSR; k0 <= H / 2; k0 += kCNT / kSR) {
      const int pr = freq_to_cell<256>(k0), qr = freq_to_cell<256>((H - k0) & (H - 1));
      if (pr == qr && pc > qcell) continue;
      const float2 cs0 = cs[k0];
      const float lam = fp.gamma + fp.alpha * (cs0.x + cs1.x);
      const float L00 = lam + fp.beta * cs0.x, L11 = lam + fp.beta * cs1.x, L01 = cs0.y * bs1;
      float A, Br, Bi;
      if (inverse) {
        const float idet = __fdividef(h, L00 * L11 - L01 * L01);
        A = idet * (L11 + L00); Br = idet * (L11 - L00); Bi = -2.0f * idet * L01;
      } else {
        A = h * (L00 + L11); Br = h * (L00 - L11); Bi = 2.0f * h * L01;
      }
      float2* zp = z + pr * kLDC + lc;
      float2* zq = z + qr * kLDC + lc2;
      const float2 Z = *zp, Zq = *zq;
      *zp = make_float2(A * Z.x + Br * Zq.x + Bi * Zq.y, A * Z.y + Bi * Zq.x - Br * Zq.y);
      if (zp != zq) *zq = make_float2(A * Zq.x + Br * Z.x + Bi * Z.y, A * Zq.y + Bi * Z.x - Br * Z.y);
    }
    __syncthreads();
  }
  fft_lines<256, kSR, +1, kCNT, kLDC, 1>(z, tw, tid);
  for (int i = tid; i < H * kSR; i += kCNT) {
    const int r = i / kSR;
    Zs[zs_row(r) + pc] = z[r * kLDC + lc];
  }
  cluster.sync();   // release/acquire at cluster scope: global writes of the other CTAs are visible
  for (int k = 0; k < NBc; ++k) {
    const int lr = k * RBc + br;
    z[lr * kLDR + c] = Zs[zs_row(r0 + lr) + c];
  }
  __syncthreads();
  fft_lines<256, kSR, +1, kCNT, 1, kLDR>(z, tw, tid);
}

#else
// Same operator with the row <-> column exchange over DISTRIBUTED SHARED MEMORY and no staging buffer: the first radix
// pass of the column FFT pulls its 16 operands per thread straight out of the four CTAs' row-layout slabs
// (ld.shared::cluster: rows n2 + 16 n1 of column pc live in CTA n1 / 4), and the first radix pass of the inverse row
// FFT pulls its 16 operands (the cells of block k1 of its row) out of the column-layout slab of the CTA that owns that
// block.  Each pull sits between two cluster barriers: one before (the producers' pass is complete everywhere), one
// after the butterflies (every CTA has its operands in registers: the slabs may be overwritten in the new layout).
// Against the L2 exchange this removes, per exchange, a 128 KiB store + 128 KiB load per CTA through L2, two
// shared-memory passes and the wait for the stores inside the barrier's release fence; it adds one cluster barrier.
// MEASURED AND REJECTED (parity green, 784 pairs): forward 16.55 ms against 11.87 ms with the L2 exchange, training step
// 43.2 against 34.0 ms - 8-byte pulls over the SM-to-SM network run at a few bytes per clock per SM, far below what
// the same data achieves as coalesced 128-byte lines through L2.  Kept as a compile-time option for the record.
template <bool inverse>
__device__ __forceinline__ void cluster_fluid(cg::cluster_group& cluster, float2* __restrict__ z, const float2* tw,
                                              const float2* cs, float2* /*Zs: unused*/, const FluidParams fp, const int tid,
                                              const int rk, const int r0, const int c, const int br, const int lc,
                                              const int q, const int pc) {
  constexpr int H = kCH, W = kCW, N1 = 16, N2 = 16, per = kSR / 16;
  static_assert(kSR * N2 == kCNT && kSR * N1 == kCNT, "one radix-16 task per thread in both pulled passes");
  (void)c; (void)br;
  fft_lines<256, kSR, -1, kCNT, 1, kLDR>(z, tw, tid);
  const unsigned zb = smem_u32(z);
  cluster.sync();                                   // row FFTs complete in all slabs
  {
    const int n2 = tid / kSR;                       // column lc (cell pc), input rows n2 + 16 n1
    float2 x[N1];
#pragma unroll
    for (int g = 0; g < kCL; ++g) {
      const unsigned rb = dsmem_of_rank(zb, g) + (unsigned)((n2 * kLDR + pc) * (int)sizeof(float2));
#pragma unroll
      for (int j = 0; j < per; ++j) x[g * per + j] = ld_dsmem2(rb + (unsigned)(16 * j * kLDR * (int)sizeof(float2)));
    }
    DftReg<N1, -1>::run(x);
#pragma unroll
    for (int k1 = 1; k1 < N1; ++k1) {
      const float2 w = tw[n2 * k1];
      x[k1] = cmul(x[k1], w.x, -w.y);
    }
    cluster.sync();                                 // every CTA holds its operands: the slab is free for the column layout
    float2* base = z + lc + n2 * kLDC;
#pragma unroll
    for (int k1 = 0; k1 < N1; ++k1) base[(N2 * k1) * kLDC] = x[k1];
  }
  __syncthreads();
  fft_fwd_pass2<256, kSR, kCNT, kLDC, 1>(z, tid);
  {
    const int k1 = cell_to_freq<256>(pc);
    const int qcell = freq_to_cell<256>((W - k1) & (W - 1));
    const int lc2 = mirror_q(rk, q) * 16 + (qcell & 15);
    const float2 cs1 = cs[k1];
    const float h = 0.5f * fp.scale, bs1 = fp.beta * cs1.y;
    for (int k0 = tid / kSR; k0 <= H / 2; k0 += kCNT / kSR) {
      const int pr = freq_to_cell<256>(k0), qr = freq_to_cell<256>((H - k0) & (H - 1));
      if (pr == qr && pc > qcell) continue;
      const float2 cs0 = cs[k0];
      const float lam = fp.gamma + fp.alpha * (cs0.x + cs1.x);
      const float L00 = lam + fp.beta * cs0.x, L11 = lam + fp.beta * cs1.x, L01 = cs0.y * bs1;
      float A, Br, Bi;
      if (inverse) {
        const float idet = __fdividef(h, L00 * L11 - L01 * L01);
        A = idet * (L11 + L00); Br = idet * (L11 - L00); Bi = -2.0f * idet * L01;
      } else {
        A = h * (L00 + L11); Br = h * (L00 - L11); Bi = 2.0f * h * L01;
      }
      float2* zp = z + pr * kLDC + lc;
      float2* zq = z + qr * kLDC + lc2;
      const float2 Z = *zp, Zq = *zq;
      *zp = make_float2(A * Z.x + Br * Zq.x + Bi * Zq.y, A * Z.y + Bi * Zq.x - Br * Zq.y);
      if (zp != zq) *zq = make_float2(A * Zq.x + Br * Z.x + Bi * Z.y, A * Zq.y + Bi * Z.x - Br * Z.y);
    }
    __syncthreads();
  }
  fft_lines<256, kSR, +1, kCNT, kLDC, 1>(z, tw, tid);
  cluster.sync();                                   // inverse column FFTs complete in all column groups
  {
    const int line = tid % kSR, k1 = tid / kSR;     // row r0 + line, cells 16 k1 .. 16 k1 + 15 = block k1
    int hg, hq;
    block_home(k1, hg, hq);
    const unsigned rb = dsmem_of_rank(zb, hg) + (unsigned)(((r0 + line) * kLDC + hq * 16) * (int)sizeof(float2));
    float2 y[N2];
#pragma unroll
    for (int k2 = 0; k2 < N2; ++k2) y[k2] = ld_dsmem2(rb + (unsigned)(k2 * (int)sizeof(float2)));
    DftReg<N2, +1>::run(y);
#pragma unroll
    for (int n2 = 1; n2 < N2; ++n2) {
      const float2 w = tw[n2 * k1];
      y[n2] = cmul(y[n2], w.x, w.y);
    }
    cluster.sync();                                 // every CTA holds its operands: the slab is free for the row layout
    float2* base = z + line * kLDR + N2 * k1;
#pragma unroll
    for (int n2 = 0; n2 < N2; ++n2) base[n2] = y[n2];
  }
  __syncthreads();
  fft_inv_pass2<256, kSR, kCNT, 1, kLDR>(z, tid);
}
#endif

template <int BG, bool LOSS>
__global__ void __launch_bounds__(kCNT, kCtasPerSM)
shoot_cluster_kernel(const ClusterParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rk = (int)cluster.block_rank();            // slab index == column-group index
  const int64_t ncl = gridDim.x / kCL;
  constexpr int H = kCH, W = kCW, N = kCN;
  float2* z = reinterpret_cast<float2*>(smem_raw);
  float2* tw = z + (size_t)H * kLDC;                   // 256 x 65 >= 64 x 257
  float2* cs = tw + 256;
  const b2_shoot_args& a = prm.a;
  const int tid = threadIdx.x;
  const int S = a.num_steps;
  const float mdt = -a.T / (float)S;
  const FluidParams fp{a.alpha, a.beta, a.gamma, 1.0f / (float)N};
  init_twiddles<256>(tw, tid, kCNT);
  init_symbol_lut<256>(cs, tid, kCNT);
  // pixel phases: a thread keeps one column and walks 16 rows of the slab
  const int c = tid % W, br = tid / W;                 // br in [0,4)
  constexpr int RBc = kCNT / W, NBc = kSR / RBc;       // 4 rows per band, 16 bands
  const int r0 = rk * kSR;
  // column phases: a thread keeps one local column of the group
  const int lc = tid % kSR, q = lc / 16, cc = lc % 16;
  const int pc = c_group_block[rk][q] * 16 + cc;        // global spectrum cell column
  __syncthreads();

  for (int64_t p = prm.p0 + blockIdx.x / kCL; p < prm.p1; p += ncl) {
    float* base = prm.scratch + (size_t)(blockIdx.x / kCL) * prm.cluster_stride;
    float* ubuf0 = base;
    float* ubuf1 = ubuf0 + 2 * (size_t)N;
    float* m0s = ubuf1 + 2 * (size_t)N;
    unsigned long long* bins = reinterpret_cast<unsigned long long*>(m0s + 2 * (size_t)N);   // sums[n], then counts
    const int64_t b = p / a.T1;
    const int t = (int)(p % a.T1);
    const float* v0p = a.v0 + (size_t)p * 2 * N;
    // The fields the gathers re-read every step (m0, u_s) stay in the per-cluster scratch (fixed addresses, L2-resident);
    // the caller's output tensors are write-only streams.
    float* m0out = (a.m0 && !a.v0_is_momentum) ? a.m0 + (size_t)p * 2 * N : nullptr;
    const float* m0r = m0s;
    float* uout = a.u + (size_t)p * 2 * N;
    float* lpart = reinterpret_cast<float*>(bins) + 3 * kMaxSectors;   // kCL x {sq, vm} partials of the loss epilogue


    // ---- load v0 (or m0) slab, m0 = flat(v0)
    for (int k = 0; k < NBc; ++k) {
      const int lr = k * RBc + br, i = (r0 + lr) * W + c;
      const float2 v = make_float2(__ldg(v0p + i), __ldg(v0p + N + i));
      z[lr * kLDR + c] = v;
      if (a.v0_is_momentum) { m0s[i] = v.x; m0s[N + i] = v.y; }
    }
    __syncthreads();
    if (!a.v0_is_momentum) {
      cluster_fluid<false>(cluster, z, tw, cs, reinterpret_cast<float2*>(ubuf0), fp, tid, rk, r0, c, br, lc, q, pc);
      for (int k = 0; k < NBc; ++k) {
        const int lr = k * RBc + br, i = (r0 + lr) * W + c;
        const float2 v = z[lr * kLDR + c];
        m0s[i] = v.x;
        m0s[N + i] = v.y;
        if (m0out) { m0out[i] = v.x; m0out[N + i] = v.y; }
      }
      __syncthreads();
    }

    const float* ucur = nullptr;
    for (int s = 0; s < S; ++s) {
      if (s > 0) {
        // m = Ad*_{u_s} m0 on the slab; u_s and m0 come from the per-pair L2 scratch (neighbour slabs included)
        const float* u0 = ucur;
        const float* u1 = ucur + N;
#ifndef B2_CLUSTER_ADSTAR_ROWS
#define B2_CLUSTER_ADSTAR_ROWS 0      // measured neutral at 256x256 (11.17 vs 11.14 ms per 784-pair shard): off
#endif
#if B2_CLUSTER_ADSTAR_ROWS
        // a thread walks NBc CONSECUTIVE rows of its column: u_s(r-1), u_s(r), u_s(r+1) slide through registers - six
        // loads per pixel (row below, left, right; both planes) instead of ten.  Same values, same arithmetic.
        {
          int clo, chi; float sc;
          diff_idx(c, W, clo, chi, sc);
          const int rfirst = r0 + br * NBc;
          float ua_up = u0[max(rfirst - 1, 0) * W + c], ub_up = u1[max(rfirst - 1, 0) * W + c];
          float ua_c = u0[rfirst * W + c], ub_c = u1[rfirst * W + c];
#pragma unroll 2
          for (int k = 0; k < NBc; ++k) {
            const int lr = br * NBc + k, r = r0 + lr;
            const int rd = min(r + 1, H - 1);
            const float sr = diff_scale(r, H);
            const float ua_dn = u0[rd * W + c], ub_dn = u1[rd * W + c];
            const float d00 = sr * (ua_dn - ua_up), d10 = sr * (ub_dn - ub_up);
            const float d01 = sc * (u0[r * W + chi] - u0[r * W + clo]), d11 = sc * (u1[r * W + chi] - u1[r * W + clo]);
            float w0, w1;
            gather2<BG, false>(m0r, N, (float)r + ua_c, (float)c + ub_c, H, W, w0, w1);
            z[lr * kLDR + c] = make_float2(w0 + (d00 * w0 + d10 * w1), w1 + (d01 * w0 + d11 * w1));
            ua_up = ua_c; ub_up = ub_c; ua_c = ua_dn; ub_c = ub_dn;
          }
        }
#else
#pragma unroll 2
        for (int k = 0; k < NBc; ++k) {
          const int lr = k * RBc + br, r = r0 + lr, i = r * W + c;
          int rlo, rhi, clo, chi; float sr, sc;
          diff_idx(r, H, rlo, rhi, sr);
          diff_idx(c, W, clo, chi, sc);
          const float d00 = sr * (u0[rhi * W + c] - u0[rlo * W + c]), d10 = sr * (u1[rhi * W + c] - u1[rlo * W + c]);
          const float d01 = sc * (u0[r * W + chi] - u0[r * W + clo]), d11 = sc * (u1[r * W + chi] - u1[r * W + clo]);
          float w0, w1;
          gather2<BG, false>(m0r, N, (float)r + u0[i], (float)c + u1[i], H, W, w0, w1);
          z[lr * kLDR + c] = make_float2(w0 + (d00 * w0 + d10 * w1), w1 + (d01 * w0 + d11 * w1));
        }
#endif
        __syncthreads();
      }
      float* unext = (s + 1 == S) ? uout : ((s & 1) ? ubuf1 : ubuf0);
      if (a.traj && s + 1 < S) unext = a.traj + ((size_t)((s + 1) * 2 + 0) * prm.P + p) * 2 * N;
      // v = sharp(m), slab in z; the spectrum is exchanged through the memory of u_{s+1} (see zs_row)
      cluster_fluid<true>(cluster, z, tw, cs, reinterpret_cast<float2*>(unext), fp, tid, rk, r0, c, br, lc, q, pc);
      float* vtraj = a.traj ? a.traj + ((size_t)(s * 2 + 1) * prm.P + p) * 2 * N : nullptr;
      float* utraj0 = (a.traj && s == 0) ? a.traj + (size_t)p * 2 * N : nullptr;
      float* velout = (s == 0 && a.vel) ? a.vel + (size_t)p * 2 * N : nullptr;
#pragma unroll 2
      for (int k = 0; k < NBc; ++k) {
        const int lr = k * RBc + br, r = r0 + lr, i = r * W + c;
        const float2 v = z[lr * kLDR + c];
        float n0 = mdt * v.x, n1 = mdt * v.y;
        if (s > 0) {
          float g0, g1;
          gather2<BG, false>(ucur, N, (float)r + n0, (float)c + n1, H, W, g0, g1);
          n0 += g0;
          n1 += g1;
        }
        unext[i] = n0;
        unext[N + i] = n1;
        if (utraj0) { utraj0[i] = 0.f; utraj0[N + i] = 0.f; }
        if (velout) { __stcs(velout + i, v.x); __stcs(velout + N + i, v.y); }      // write-only outputs: streaming stores
        if (vtraj) { __stcs(vtraj + i, v.x); __stcs(vtraj + N + i, v.y); }
      }
      if (LOSS && s == 0) {
        // loss epilogue, regularisation term of this slab: v_0 . m0 (v_0 = vel just stored); reduced here so that no
        // accumulator lives across the geodesic.  z still holds v and is not written before the next barrier.
        float acc_vm = 0.f, none = 0.f;
        for (int k = 0; k < NBc; ++k) {
          const int lr = k * RBc + br, i = (r0 + lr) * W + c;
          const float2 v = z[lr * kLDR + c];
          acc_vm += v.x * m0r[i] + v.y * m0r[N + i];
        }
        block_reduce2<kCNT>(acc_vm, none, reinterpret_cast<float*>(cs + 256), tid);
        if (tid == 0) lpart[2 * rk + 1] = acc_vm;
      }
      ucur = unext;
      cluster.sync();   // release/acquire at cluster scope: global writes of the other CTAs are visible                                  // u_{s+1} of all four slabs visible to the cluster
    }

    // ---- deformed source on the slab
    if (a.sdef || LOSS) {
      const float* src = a.src_per_pair
                             ? (a.src_slice_stride ? a.src + (size_t)b * a.src_slice_stride + (size_t)t * N
                                                   : a.src + (size_t)p * N)
                             : a.src + (size_t)b * (a.src_slice_stride ? a.src_slice_stride : N);
      float* sd = a.sdef ? a.sdef + (size_t)p * N : nullptr;
      const float* tarp = nullptr;
      if (LOSS)
        tarp = a.tar_slice_stride ? a.tar + (size_t)b * a.tar_slice_stride + (size_t)t * N : a.tar + (size_t)p * N;
      float acc_sq = 0.f;
      for (int k = 0; k < NBc; ++k) {
        const int lr = k * RBc + br, r = r0 + lr, i = r * W + c;
        const float val = gather1_ldg<BG>(src, (float)r + ucur[i], (float)c + ucur[N + i], H, W);
        if (sd) __stcs(sd + i, val);
        if (LOSS) {
          const float d = __ldg(tarp + i) - val;
          acc_sq += d * d;
        }
      }
      if (LOSS) {
        // slab partials in fixed order; rank 0 adds the four after the closing cluster barrier
        float none = 0.f;
        block_reduce2<kCNT>(acc_sq, none, reinterpret_cast<float*>(cs + 256), tid);
        if (tid == 0) lpart[2 * rk] = acc_sq;
      }
    }

    // ---- strain: every CTA bins its slab into the per-pair bins in L2 (integer atomics), rank 0 writes the column
    if (a.S) {
      const int ns = a.n_sectors;
      unsigned int* cnts = reinterpret_cast<unsigned int*>(bins + ns);
      if (rk == 0)
        for (int i = tid; i < ns; i += kCNT) { bins[i] = 0ull; cnts[i] = 0u; }
      cluster.sync();   // release/acquire at cluster scope: global writes of the other CTAs are visible
      const long long* mom = reinterpret_cast<const long long*>(a.moments) + 3 * b;
      const SectorFrame sf = sector_frame_of(a.table, a.table_slice_stride, a.theta0, a.clockwise, b);
      const long long cnt = mom[0], sx = mom[1], sy = mom[2];
      float c0, c1;
      centroid_from_moments(mom, H, W, c0, c1);
      const float* tarp = a.tar_slice_stride ? a.tar + (size_t)b * a.tar_slice_stride + (size_t)t * N
                                             : a.tar + (size_t)p * N;
      const float* u0 = ucur;
      const float* u1 = ucur + N;
#ifndef B2_CLUSTER_STRAIN_COMPACT
#define B2_CLUSTER_STRAIN_COMPACT 1
#endif
#if B2_CLUSTER_STRAIN_COMPACT
      // member pixels of the slab compacted first (as in shoot_fwd_kernel: strain_bin_frame_compact): the slab buffer z
      // is dead after the last compose and holds the list of slab-local pixel indices (64 x 256 = 16 bits)
      unsigned short* list = reinterpret_cast<unsigned short*>(z);
      __shared__ int n_mem_s;
      if (tid == 0) n_mem_s = 0;
      __syncthreads();
      for (int k = 0; k < NBc; ++k) {
        const int lr = k * RBc + br;
        const bool mem = tarp[(r0 + lr) * W + c] > 0.5f;
        const unsigned bal = __ballot_sync(0xffffffffu, mem);
        if (bal) {
          int base = 0;
          if ((tid & 31) == 0) base = atomicAdd(&n_mem_s, __popc(bal));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (mem) list[base + __popc(bal & ((1u << (tid & 31)) - 1u))] = (unsigned short)(lr * W + c);
        }
      }
      __syncthreads();
      const int n_mem = n_mem_s;
      for (int j = tid; j < n_mem; j += kCNT) {
        const int li = list[j], cm = li % W, r = r0 + li / W, i = r * W + cm;
        const int ksec = classify_sector(cnt * r - sx, cnt * cm - sy, sf.table, ns, sf.theta0, sf.flip);
        if (ksec < 0) continue;
        int rlo, rhi, clo, chi; float sr, sc;
        diff_idx(r, H, rlo, rhi, sr);
        diff_idx(cm, W, clo, chi, sc);
        const float d00 = sr * (u0[rhi * W + cm] - u0[rlo * W + cm]), d10 = sr * (u1[rhi * W + cm] - u1[rlo * W + cm]);
        const float d01 = sc * (u0[r * W + chi] - u0[r * W + clo]), d11 = sc * (u1[r * W + chi] - u1[r * W + clo]);
        EccTerms e; float ecc;
        if (!ecc_eval(d00, d01, d10, d11, (float)r + u0[i], (float)cm + u1[i], c0, c1, e, ecc)) continue;
        atomicAdd(bins + ksec, ecc_to_fixed(ecc));
        atomicAdd(cnts + ksec, 1u);
      }
#else
      for (int k = 0; k < NBc; ++k) {
        const int lr = k * RBc + br, r = r0 + lr, i = r * W + c;
        if (!(tarp[i] > 0.5f)) continue;
        const int ksec = classify_sector(cnt * r - sx, cnt * c - sy, sf.table, ns, sf.theta0, sf.flip);
        if (ksec < 0) continue;
        int rlo, rhi, clo, chi; float sr, sc;
        diff_idx(r, H, rlo, rhi, sr);
        diff_idx(c, W, clo, chi, sc);
        const float d00 = sr * (u0[rhi * W + c] - u0[rlo * W + c]), d10 = sr * (u1[rhi * W + c] - u1[rlo * W + c]);
        const float d01 = sc * (u0[r * W + chi] - u0[r * W + clo]), d11 = sc * (u1[r * W + chi] - u1[r * W + clo]);
        EccTerms e; float ecc;
        if (!ecc_eval(d00, d01, d10, d11, (float)r + u0[i], (float)c + u1[i], c0, c1, e, ecc)) continue;
        atomicAdd(bins + ksec, ecc_to_fixed(ecc));
        atomicAdd(cnts + ksec, 1u);
      }
#endif
      cluster.sync();   // release/acquire at cluster scope: global writes of the other CTAs are visible
      if (rk == 0) {
        for (int k = tid; k < ns; k += kCNT) {
          const int cn = (int)__ldcg(cnts + k);
          const float val = fixed_to_mean(__ldcg(bins + k), cn);
          float* row = a.S + ((size_t)b * ns + k) * a.n_frames;
          if (t < a.n_frames) row[t] = val;
          if (t == (int)a.T1 - 1)
            for (int tt = (int)a.T1; tt < a.n_frames; ++tt) row[tt] = val;
          if (a.counts) a.counts[((size_t)b * ns + k) * a.T1 + t] = cn;
        }
      }
    }
    cluster.sync();                                     // scratch free for the next pair of this cluster
    if (LOSS && rk == 0 && tid == 0) {
      float sq = 0.f, vm = 0.f;                      // slab partials in fixed order: bitwise reproducible
      for (int x = 0; x < kCL; ++x) { sq += __ldcg(lpart + 2 * x); vm += __ldcg(lpart + 2 * x + 1); }
      a.loss_terms[2 * p] = sq;
      a.loss_terms[2 * p + 1] = vm;
    }
  }
}

constexpr size_t kClusterSmem = sizeof(float2) * ((size_t)kCH * kLDC + 512 + 32);   // z + twiddles + symbol LUT + reduction scratch

static size_t cluster_scratch_floats() { return (size_t)6 * kCN + 4 * kMaxSectors; }   // u(2x2) + m0(2) fields + bins

template <int BG>
static int cluster_max_active(int* out) {
  B2_CUDA(cudaFuncSetAttribute(shoot_cluster_kernel<BG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClusterSmem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kCL * 64);
  cfg.blockDim = dim3(kCNT);
  cfg.dynamicSmemBytes = kClusterSmem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  B2_CUDA(cudaOccupancyMaxActiveClusters(&n, shoot_cluster_kernel<BG, false>, &cfg));
  *out = n;
  return B2_OK;
}

// number of clusters to launch (co-resident clusters, at most P); 0 when clusters are unavailable
int cluster_grid_clusters(int64_t P) {
  static int cached = -1;
  if (cached < 0) {
    int n = 0;
    if (cluster_max_active<B2_BG_CLAMP>(&n) != B2_OK || n < 1) { (void)cudaGetLastError(); return 0; }
    cached = n;
  }
  return (int)(cached < P ? cached : P);
}

int64_t cluster_workspace_bytes(int64_t P) {
  int n = cluster_grid_clusters(P);
  if (n < 1) n = 37 < P ? 37 : (int)P;                 // no device (size query on a CPU box): assume a full B200
  return (int64_t)(sizeof(float) * cluster_scratch_floats() * (size_t)n);
}

int launch_shoot_cluster(const b2_shoot_args& a, void* workspace, cudaStream_t st) {
  const int64_t P = a.B * a.T1;
  const int64_t p0 = a.pair_count > 0 ? a.pair_begin : 0, np = a.pair_count > 0 ? a.pair_count : P;
  const int ncl = cluster_grid_clusters(np);
  if (ncl < 1) return B2_E_FFTSIZE;
  ClusterParams prm;
  prm.a = a;
  prm.scratch = reinterpret_cast<float*>(workspace);
  prm.P = P;
  prm.p0 = p0;
  prm.p1 = p0 + np;
  prm.cluster_stride = (int64_t)cluster_scratch_floats();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(kCL * ncl));
  cfg.blockDim = dim3(kCNT);
  cfg.dynamicSmemBytes = kClusterSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#define B2_CLUSTER_LAUNCH(BGV, LV)                                                                              \
  do {                                                                                                          \
    B2_CUDA(cudaFuncSetAttribute(shoot_cluster_kernel<BGV, LV>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                 (int)kClusterSmem));                                                           \
    B2_CUDA(cudaLaunchKernelEx(&cfg, shoot_cluster_kernel<BGV, LV>, prm));                                      \
  } while (0)
  const bool loss = a.loss_terms != nullptr;   // compile-time variant: the inference kernel carries no epilogue code
  if (a.background == B2_BG_CLAMP) {
    if (loss) B2_CLUSTER_LAUNCH(B2_BG_CLAMP, true); else B2_CLUSTER_LAUNCH(B2_BG_CLAMP, false);
  } else {
    if (loss) B2_CLUSTER_LAUNCH(B2_BG_ZERO, true); else B2_CLUSTER_LAUNCH(B2_BG_ZERO, false);
  }
#undef B2_CLUSTER_LAUNCH
  return B2_OK;
}

// ------------------------------------------------------------------ fused EPDiff adjoint at 256x256
// Same reverse sweep as shoot_bwd_kernel (shoot.cu), one 4-CTA cluster per frame-pair, 64-row slabs.  dL/dv_s -> dL/dm_s
// lives in the slab in shared memory (cluster_fluid is the self-adjoint sharp); the accumulators dL/du (ping-pong),
// dL/dm0 and w = m0 o (id + u_s) sit in a per-cluster scratch that stays in L2 (4 fields = 2 MiB per cluster).  Every
// scratch access is on the L2 path (ld.global.cg / st.global.cg / RED), so no stale L1 line can be observed across SMs.
// Splats cross slab boundaries (float REDs into the shared scratch, pre-aggregated by splat2_agg); the only other
// cross-slab data are the row-neighbour products g_0 w_b of the Ad* adjoint (L2).  The spectrum exchange of the fluid
// operator reuses the consumed dL/du_{s+1} buffer (each CTA writes only its own slab rows of it, see zs_row).
// Four cluster barriers per adjoint step.
template <int BG>
__global__ void __launch_bounds__(kCNT, kCtasPerSM)
shoot_cluster_bwd_kernel(const ShootBwdParams prm, const int64_t cluster_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rk = (int)cluster.block_rank();
  const int64_t ncl = gridDim.x / kCL;
  constexpr int H = kCH, W = kCW, N = kCN;
  float2* z = reinterpret_cast<float2*>(smem_raw);
  float2* tw = z + (size_t)H * kLDC;
  float2* cs = tw + 256;
  const int tid = threadIdx.x, lane = tid & 31;
  const int S = prm.num_steps;
  const float mdt = -prm.T / (float)S;
  const FluidParams fp{prm.alpha, prm.beta, prm.gamma, 1.0f / (float)N};
  init_twiddles<256>(tw, tid, kCNT);
  init_symbol_lut<256>(cs, tid, kCNT);
  const int c = tid % W, br = tid / W;                 // cluster_fluid's slab <-> scratch map (rows strided by 4)
  constexpr int NBc = kSR / (kCNT / W);                // 16 pixels per thread
  const int r0 = rk * kSR;
  const int lc = tid % kSR, q = lc / 16, cc = lc % 16;
  const int pc = c_group_block[rk][q] * 16 + cc;
  // pixel phases: a thread keeps one column and walks 16 CONSECUTIVE rows of the slab (splat2_agg carries the
  // bottom-row contributions from row to row)
  const int lrbase = br * NBc;
  const int cl = max(c - 1, 0), cr = min(c + 1, W - 1);
  const float sc = diff_scale(c, W);
  const float cmc = (c >= 1) ? diff_scale(c - 1, W) : 0.f, cpc = (c <= W - 2) ? diff_scale(c + 1, W) : 0.f;
  const float c0c = (c == W - 1 ? 1.f : 0.f) - (c == 0 ? 1.f : 0.f);
  float* base = prm.scratch + (size_t)(blockIdx.x / kCL) * cluster_stride;
  float* Ga = base;
  float* Gb = Ga + 2 * (size_t)N;
  float* A = Gb + 2 * (size_t)N;
  float* Qb = A + 2 * (size_t)N;                       // (g_0 w_0, g_0 w_1): row-neighbour products of the Ad* adjoint
  __syncthreads();

  for (int64_t p = blockIdx.x / kCL; p < prm.P; p += ncl) {
    const float* m0p = prm.m0 + (size_t)p * 2 * N;
    float* Gcur = Ga;
    float* Gnext = Gb;
    for (int k = 0; k < NBc; ++k) {
      const int i = (r0 + lrbase + k) * W + c;
      __stcg(Gcur + i, prm.gu ? __ldg(prm.gu + (size_t)p * 2 * N + i) : 0.f);
      __stcg(Gcur + N + i, prm.gu ? __ldg(prm.gu + (size_t)p * 2 * N + N + i) : 0.f);
      __stcg(A + i, prm.gm0 ? __ldg(prm.gm0 + (size_t)p * 2 * N + i) : 0.f);
      __stcg(A + N + i, prm.gm0 ? __ldg(prm.gm0 + (size_t)p * 2 * N + N + i) : 0.f);
      __stcg(Gnext + i, 0.f);
      __stcg(Gnext + N + i, 0.f);
    }
    cluster.sync();

    for (int s = S - 1; s >= 0; --s) {
      const float* us = prm.traj + ((size_t)(2 * s) * prm.P + p) * 2 * N;
      const float* vs = prm.traj + ((size_t)(2 * s + 1) * prm.P + p) * 2 * N;
      if (s > 0) {
        // ---- adjoint of u_{s+1} = interp(u_s, v_s, -dt) - dt v_s : dL/dv_s -> z, splat of dL/du_{s+1} -> Gnext
        SplatCarry cy{-1, 0.f, 0.f};
#pragma unroll 2
        for (int k = 0; k < NBc; ++k) {
          const int lr = lrbase + k, r = r0 + lr, i = r * W + c;
          const float g0 = __ldcg(Gcur + i), g1 = __ldcg(Gcur + N + i);
          const float v0 = __ldg(vs + i), v1 = __ldg(vs + N + i);
          const Taps t = make_taps<BG>((float)r + mdt * v0, (float)c + mdt * v1, H, W);
          float a0, a1, b0, b1;
          tap_grad<BG>(t, __ldg(us + t.o00), __ldg(us + t.o10), __ldg(us + t.o01), __ldg(us + t.o11), a0, a1);
          tap_grad<BG>(t, __ldg(us + N + t.o00), __ldg(us + N + t.o10), __ldg(us + N + t.o01), __ldg(us + N + t.o11), b0, b1);
          z[lr * kLDR + c] = make_float2(mdt * (g0 * a0 + g1 * b0 + g0), mdt * (g0 * a1 + g1 * b1 + g1));
          splat2_agg<BG>(Gnext, N, t, g0, g1, cy, lane);
        }
        splat_flush(Gnext, N, cy);
      } else {
        // u_0 = 0: u_1 = -dt v_0, so dL/dv_0 = -dt dL/du_1 (+ the direct gradient of the velocity output)
        const float* gv = prm.gvel ? prm.gvel + (size_t)p * 2 * N : nullptr;
        for (int k = 0; k < NBc; ++k) {
          const int lr = lrbase + k, i = (r0 + lr) * W + c;
          float a = mdt * __ldcg(Gcur + i), b = mdt * __ldcg(Gcur + N + i);
          if (gv) { a += __ldg(gv + i); b += __ldg(gv + N + i); }
          z[lr * kLDR + c] = make_float2(a, b);
        }
      }
      __syncthreads();
      // ---- dL/dm_s = sharp(dL/dv_s); the spectrum is exchanged through the consumed dL/du_{s+1} buffer
      cluster_fluid<true>(cluster, z, tw, cs, reinterpret_cast<float2*>(Gcur), fp, tid, rk, r0, c, br, lc, q, pc);
      if (s > 0) {
        // ---- adjoint of m_s = (I + Du_s)^T (m0 o (id + u_s)) in two passes, as in shoot_bwd_kernel: pass A does
        // everything local to a pixel with one gather of m0 and leaves the products g_a w_b behind - (g_1 w_0,
        // g_1 w_1) in place of g in the slab (column neighbours never leave the slab), (g_0 w_0, g_0 w_1) in the
        // scratch field Qb (row neighbours cross slab boundaries: L2) - pass B adds their transposed differences.
        SplatCarry cy{-1, 0.f, 0.f};
#ifndef B2_CLUSTER_BWD_ROWWIN
#define B2_CLUSTER_BWD_ROWWIN 1   // u_s rows and row-neighbour products slide through registers (as in shoot_bwd_kernel)
#endif
#if B2_CLUSTER_BWD_ROWWIN
        const int rfirst = r0 + lrbase;
        float ua_up = __ldg(us + max(rfirst - 1, 0) * W + c), ub_up = __ldg(us + N + max(rfirst - 1, 0) * W + c);
        float ua_c = __ldg(us + rfirst * W + c), ub_c = __ldg(us + N + rfirst * W + c);
#endif
#pragma unroll 2
        for (int k = 0; k < NBc; ++k) {
          const int lr = lrbase + k, r = r0 + lr, i = r * W + c;
          const float gn0 = __ldcg(Gnext + i), gn1 = __ldcg(Gnext + N + i);
          const int ru = max(r - 1, 0), rd = min(r + 1, H - 1);
          const int oup = ru * W + c, odn = rd * W + c, olf = r * W + cl, ort = r * W + cr;
          const float sr = diff_scale(r, H);
#if B2_CLUSTER_BWD_ROWWIN
          const float ua_dn = __ldg(us + odn), ub_dn = __ldg(us + N + odn);
          const float d00 = sr * (ua_dn - ua_up), d10 = sr * (ub_dn - ub_up);
          const float uc0 = ua_c, uc1 = ub_c;
          ua_up = ua_c; ub_up = ub_c; ua_c = ua_dn; ub_c = ub_dn;
          (void)oup;
#else
          const float d00 = sr * (__ldg(us + odn) - __ldg(us + oup)), d10 = sr * (__ldg(us + N + odn) - __ldg(us + N + oup));
          const float uc0 = __ldg(us + i), uc1 = __ldg(us + N + i);
#endif
          const float d01 = sc * (__ldg(us + ort) - __ldg(us + olf)), d11 = sc * (__ldg(us + N + ort) - __ldg(us + N + olf));
          const float2 g = z[lr * kLDR + c];
          const float gw0 = g.x + (d00 * g.x + d01 * g.y);
          const float gw1 = g.y + (d10 * g.x + d11 * g.y);
          const Taps t = make_taps<BG>((float)r + uc0, (float)c + uc1, H, W);
          splat2_agg<BG>(A, N, t, gw0, gw1, cy, lane);
          float w0, w1, o0, o1;
          {
            const float v00 = __ldg(m0p + t.o00), v10 = __ldg(m0p + t.o10), v01 = __ldg(m0p + t.o01), v11 = __ldg(m0p + t.o11);
            float a0, a1;
            w0 = tap_sample<BG>(t, v00, v10, v01, v11);
            tap_grad<BG>(t, v00, v10, v01, v11, a0, a1);
            o0 = gw0 * a0;
            o1 = gw0 * a1;
          }
          {
            const float v00 = __ldg(m0p + N + t.o00), v10 = __ldg(m0p + N + t.o10), v01 = __ldg(m0p + N + t.o01),
                        v11 = __ldg(m0p + N + t.o11);
            float b0, b1;
            w1 = tap_sample<BG>(t, v00, v10, v01, v11);
            tap_grad<BG>(t, v00, v10, v01, v11, b0, b1);
            o0 += gw1 * b0;
            o1 += gw1 * b1;
          }
          __stcg(Gnext + i, gn0 + o0);          // own pixel: the compose adjoint's REDs completed two barriers ago
          __stcg(Gnext + N + i, gn1 + o1);
          __stcg(Qb + i, g.x * w0);
          __stcg(Qb + N + i, g.x * w1);
          z[lr * kLDR + c] = make_float2(g.y * w0, g.y * w1);
        }
        splat_flush(A, N, cy);
        cluster.sync();                          // row-neighbour products of the adjacent slabs are visible
#if B2_CLUSTER_BWD_ROWWIN
        float qa_up = __ldcg(Qb + max(rfirst - 1, 0) * W + c), qb_up = __ldcg(Qb + N + max(rfirst - 1, 0) * W + c);
        float qa_c = __ldcg(Qb + rfirst * W + c), qb_c = __ldcg(Qb + N + rfirst * W + c);
#endif
#pragma unroll 2
        for (int k = 0; k < NBc; ++k) {
          const int lr = lrbase + k, r = r0 + lr, i = r * W + c;
          const float gn0 = __ldcg(Gnext + i), gn1 = __ldcg(Gnext + N + i);
          const int ru = max(r - 1, 0), rd = min(r + 1, H - 1);
          const int oup = ru * W + c, odn = rd * W + c;
          const float cmr = (r >= 1) ? diff_scale(r - 1, H) : 0.f, cpr = (r <= H - 2) ? diff_scale(r + 1, H) : 0.f;
          const float c0r = (r == H - 1 ? 1.f : 0.f) - (r == 0 ? 1.f : 0.f);
          const float2 ql = z[lr * kLDR + cl], qr = z[lr * kLDR + cr];
#if B2_CLUSTER_BWD_ROWWIN
          const float qa_dn = __ldcg(Qb + odn), qb_dn = __ldcg(Qb + N + odn);
          float o0 = (cmr * qa_up - cpr * qa_dn) + (cmc * ql.x - cpc * qr.x);
          float o1 = (cmr * qb_up - cpr * qb_dn) + (cmc * ql.y - cpc * qr.y);
          if (c0r != 0.f) { o0 += c0r * qa_c; o1 += c0r * qb_c; }
          qa_up = qa_c; qb_up = qb_c; qa_c = qa_dn; qb_c = qb_dn;
          (void)oup;
#else
          float o0 = (cmr * __ldcg(Qb + oup) - cpr * __ldcg(Qb + odn)) + (cmc * ql.x - cpc * qr.x);
          float o1 = (cmr * __ldcg(Qb + N + oup) - cpr * __ldcg(Qb + N + odn)) + (cmc * ql.y - cpc * qr.y);
          if (c0r != 0.f) { o0 += c0r * __ldcg(Qb + i); o1 += c0r * __ldcg(Qb + N + i); }
#endif
          if (c0c != 0.f) { const float2 qc = z[lr * kLDR + c]; o0 += c0c * qc.x; o1 += c0c * qc.y; }
          __stcg(Gnext + i, gn0 + o0);
          __stcg(Gnext + N + i, gn1 + o1);
          __stcg(Gcur + i, 0.f);                // becomes the splat target of the next step (dead since the exchange)
          __stcg(Gcur + N + i, 0.f);
        }
        float* tmp = Gcur; Gcur = Gnext; Gnext = tmp;
        cluster.sync();   // dL/dm0 REDs, product reads and the zero fill are complete cluster-wide
      } else {
        // m_0 = Ad*_0 m0 = m0 exactly: dL/dm0 = accumulated splats + dL/dm_0, summed in place in the slab
        for (int k = 0; k < NBc; ++k) {
          const int lr = lrbase + k, i = (r0 + lr) * W + c;
          float2 g = z[lr * kLDR + c];
          g.x += __ldcg(A + i);
          g.y += __ldcg(A + N + i);
          z[lr * kLDR + c] = g;
        }
        __syncthreads();
      }
    }
    // ---- dL/dv0 = flat(dL/dm0)  (or dL/dm0 itself when the forward input was the momentum)
    if (!prm.v0_is_momentum)
      cluster_fluid<false>(cluster, z, tw, cs, reinterpret_cast<float2*>(Gcur), fp, tid, rk, r0, c, br, lc, q, pc);
    float* out = prm.gv0 + (size_t)p * 2 * N;
    const float g2 = prm.g_reg ? 2.f * __ldg(prm.g_reg + p) : 0.f;
    const float* radd = prm.v0_is_momentum ? prm.traj + ((size_t)prm.P + p) * 2 * N : m0p;   // v_0 of the trajectory
    for (int k = 0; k < NBc; ++k) {
      const int lr = lrbase + k, i = (r0 + lr) * W + c;
      float2 v = z[lr * kLDR + c];
      if (prm.g_reg) { v.x += g2 * __ldg(radd + i); v.y += g2 * __ldg(radd + N + i); }
      out[i] = v.x;
      out[N + i] = v.y;
    }
    cluster.sync();   // scratch and slab free for the next pair of this cluster
  }
}

static size_t cluster_bwd_scratch_floats() { return (size_t)8 * kCN; }   // 4 fields: G ping, G pong, dL/dm0, products

template <int BG>
static int cluster_bwd_max_active(int* out) {
  B2_CUDA(cudaFuncSetAttribute(shoot_cluster_bwd_kernel<BG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClusterSmem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kCL * 64);
  cfg.blockDim = dim3(kCNT);
  cfg.dynamicSmemBytes = kClusterSmem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  B2_CUDA(cudaOccupancyMaxActiveClusters(&n, shoot_cluster_bwd_kernel<BG>, &cfg));
  *out = n;
  return B2_OK;
}

static int cluster_bwd_grid_clusters(int64_t P) {
  static int cached = -1;
  if (cached < 0) {
    int n = 0;
    if (cluster_bwd_max_active<B2_BG_CLAMP>(&n) != B2_OK || n < 1) { (void)cudaGetLastError(); return 0; }
    cached = n;
  }
  return (int)(cached < P ? cached : P);
}

int64_t cluster_bwd_workspace_bytes(int64_t P) {
  int n = cluster_bwd_grid_clusters(P);
  if (n < 1) n = 37 < P ? 37 : (int)P;                 // size query without a device: assume a full B200
  return (int64_t)(sizeof(float) * cluster_bwd_scratch_floats() * (size_t)n);
}

int launch_shoot_cluster_bwd(const ShootBwdParams& prm, int background, cudaStream_t st) {
  const int ncl = cluster_bwd_grid_clusters(prm.P);
  if (ncl < 1) return B2_E_FFTSIZE;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(kCL * ncl));
  cfg.blockDim = dim3(kCNT);
  cfg.dynamicSmemBytes = kClusterSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const int64_t stride = (int64_t)cluster_bwd_scratch_floats();
  if (background == B2_BG_CLAMP) {
    B2_CUDA(cudaFuncSetAttribute(shoot_cluster_bwd_kernel<B2_BG_CLAMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClusterSmem));
    B2_CUDA(cudaLaunchKernelEx(&cfg, shoot_cluster_bwd_kernel<B2_BG_CLAMP>, prm, stride));
  } else {
    B2_CUDA(cudaFuncSetAttribute(shoot_cluster_bwd_kernel<B2_BG_ZERO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClusterSmem));
    B2_CUDA(cudaLaunchKernelEx(&cfg, shoot_cluster_bwd_kernel<B2_BG_ZERO>, prm, stride));
  }
  return B2_OK;
}

}  // namespace b2

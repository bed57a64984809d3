// Fused geodesic shooting: flat + S x EPDiff_step + warp + strain/sector reduction.
// Replaces the body of forward_volume (call site
// /root/reference/modules/trainer/joint_registration_strainmat_LMA.py:307):
//   m0 = metric.flat(v0); u = lm.expmap(metric, m0, num_steps=S); Sdef = lm.interp(src, u);
//   strain_matrix = sector_reduce(strain(u), mask)            
//
// Path A (square grids up to 128x128): ONE persistent kernel, one CTA per
// frame-pair at a time, grid = #SMs x resident CTAs.  The complex-packed working
// field (m / v) lives in shared memory for the whole geodesic: every sharp() is
// an in-SM FFT -> symbol multiply -> IFFT with no HBM round trip.  m0 and the
// displacement u_s are written once per step to global memory and re-read by
// the same CTA for the 4-tap gathers, so they are served from L1/L2 (the
// working set of all resident CTAs is ~57 MB at 128^2, inside the 126 MB L2).
// Compulsory HBM traffic per pair drops from the op-level 700*N bytes to the
// inputs + outputs (~44*N bytes).
//
// 256x256: one 4-CTA thread-block cluster per frame-pair (shoot_cluster.cu), same residency idea with 64-row slabs.
//
// Path B (rectangular grids, B2_FLAG_OPLEVEL at 256x256, or a device that cannot co-schedule the cluster): the same
// algorithm as a sequence of the op-level kernels (HBM-bound per op).
//
// Backward: shoot_bwd_kernel (fused EPDiff adjoint, square grids up to 128x128), shoot_cluster_bwd_kernel (256x256,
// shoot_cluster.cu) or the op-level sweep; b2_shoot_bwd_ex selects.
//
// Both single-CTA kernels draw their work from an atomic ticket counter (whole pairs, then the last grid-size pairs in
// chunks of two steps handed from CTA to CTA through global memory) instead of a static round-robin: see
// "dynamic schedule" below.
#include "fft.cuh"
#include "shoot_params.cuh"
#include "strain.cuh"

namespace b2 {

int fluid_apply_impl(const float* f, float* out, int64_t P, int64_t H, int64_t W, float alpha, float beta,
                     float gamma, int inverse, void* workspace, int64_t workspace_bytes, cudaStream_t st);
int adstar_bwd_impl(const float* gout, const float* u, const float* m0, float* du, float* dm0, float* workspace,
                    int64_t P, int64_t H, int64_t W, int background, bool zero_dm0, cudaStream_t st,
                    const float* du_add);

// 256x256: one 4-CTA cluster per pair (shoot_cluster.cu)
int cluster_grid_clusters(int64_t P);
int64_t cluster_workspace_bytes(int64_t P);
int launch_shoot_cluster(const b2_shoot_args& a, void* workspace, cudaStream_t st);
int strain_sector_fwd_range(const float* u, const float* tar, const int64_t* moments, const b2_sector_frame* frame,
                            float* S, int32_t* counts, int64_t B, int64_t T1, int64_t H, int64_t W, int n_sectors,
                            int n_frames, int64_t p0, int64_t np, cudaStream_t st);   // strain.cu
int64_t cluster_bwd_workspace_bytes(int64_t P);
int launch_shoot_cluster_bwd(const ShootBwdParams& prm, int background, cudaStream_t st);
// 256x256 runs on the cluster kernels unless the caller asks for the op-level path or the device cannot
// co-schedule a 4-CTA cluster with this much shared memory (then path B serves it)
static bool cluster_size(int64_t H, int64_t W, int64_t P, int flags) {
  return H == 256 && W == 256 && !(flags & B2_FLAG_OPLEVEL) && cluster_grid_clusters(P) != 0;
}
constexpr bool kClusterBwd = true;    // fused 256x256 adjoint (shoot_cluster_bwd_kernel)
static bool cluster_bwd_size(int64_t H, int64_t W, int64_t P, int flags) { return kClusterBwd && cluster_size(H, W, P, flags); }

int epdiff_step_big(const float* u, const float* m0, float* unext, float* vout, void* zg, int64_t P, int64_t H,
                    int64_t W, float alpha, float beta, float gamma, float dt, int bg, cudaStream_t st);

constexpr int kFusedMaxSectors = 256;
static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }
#ifndef B2_FAST_GATHER
#define B2_FAST_GATHER 0
#endif
constexpr bool kFastGather = B2_FAST_GATHER != 0;
#ifndef B2_COMPOSE_UNROLL
#define B2_COMPOSE_UNROLL 2
#endif
constexpr int kComposeUnroll = B2_COMPOSE_UNROLL;

struct ShootParams {
  b2_shoot_args a;
  float* scratch;     // per-CTA: [u ping-pong | m0] fields + hand-over field of the tail pair; then ticket + flags
  int64_t P;
  int64_t field;      // 2*H*W floats
  int balanced;       // != 0: dynamic ticket schedule with hand-over of the tail pairs (below)
};

// ---- dynamic schedule of the persistent grid (removes the partial last wave) ----------------------------------
// With one CTA per pair at a time and P not a multiple of the grid, the last wave runs on a fraction of the SMs
// (configs[1]: 1536 pairs on 148 CTAs = 10.38 waves, 5.6 % of the kernel idle), and SMs do not all run at the same
// speed.  A geodesic cannot be split in space at 128x128, but it can be HANDED OVER between steps: its whole state at
// a step boundary is (u_s, m0) in global memory.  Work is therefore drawn from ONE atomic ticket counter:
//   tickets [0, P - K)            whole pairs (K = grid size: the last K pairs are the "tail"),
//   tickets P - K + c K + i       chunk c (kChunkSteps EPDiff steps) of tail pair i; chunk 0 includes the prologue,
//                                 the last chunk the epilogue.
// Chunks of one pair are K tickets apart, so by the time chunk c of a pair is drawn its chunk c-1 was drawn K tickets
// earlier (normally finished; a flag per (pair, chunk) makes it a guarantee).  The CTA that holds an earlier ticket is
// running, so waits cannot deadlock.  The kernel ends within one chunk of perfect balance instead of one pair, and
// fast SMs simply draw more tickets.  Same arithmetic per pair whoever executes it: results are bit-identical.
// v_s of the trajectory, the velocity output and the warped source are write-only in the forward kernel (the adjoint
// / the caller read them much later): streaming stores keep them from displacing the L2-resident scratch
#ifndef B2_TRAJ_STREAM
#define B2_TRAJ_STREAM 1
#endif
#if B2_TRAJ_STREAM
#define B2_TRAJ_STORE(p, v) __stcs((p), (v))
#define B2_ONCE_LOAD(p) __ldcs(p)
#else
#define B2_TRAJ_STORE(p, v) (*(p) = (v))
#define B2_ONCE_LOAD(p) __ldg(p)
#endif
#ifndef B2_BALANCED_BWD
#define B2_BALANCED_BWD 1
#endif
constexpr int kBwdMaxChunks = 64;
constexpr int kChunkSteps = 2;     // >= 2: a chunk's first step reads the hand-over buffer, its last step rewrites it
#ifndef B2_BALANCED
#define B2_BALANCED 1              // 0: static schedule (CTA j takes pairs j, j + G, ...)
#endif

template <int H, int W>
struct ShootSmem {
  using FS = FluidSmem<H, W>;
  static constexpr size_t bins_off = (FS::bytes + 15) & ~size_t(15);
  static constexpr size_t red_off = bins_off + sizeof(int32_t) * 5 * kFusedMaxSectors;   // loss-epilogue partials
  static constexpr size_t bytes = red_off + sizeof(float) * 64;
};

// Thread <-> pixel map of every per-pixel phase: a thread keeps ONE column c = tid % W and walks the
// rows r = tid / W, + RB, + 2 RB ... (RB = NT / W rows per band), so lanes run along the contiguous axis
// (coalesced global rows, conflict-free smem rows) and all addresses advance by constants.
template <int H, int W, int NT, int BG, bool LOSS>
__global__ void __launch_bounds__(NT)
shoot_fwd_kernel(const ShootParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using FS = FluidSmem<H, W>;
  static_assert(NT % W == 0 && H % (NT / W) == 0, "CTA must tile the grid in whole row bands");
  constexpr int LD = FS::LD, N = H * W, RB = NT / W, NB = H / RB;
  float2 *z, *twH, *twW, *csH, *csW;
  FS::carve(smem_raw, z, twH, twW, csH, csW);
  const b2_shoot_args& a = prm.a;
  const int n_sectors = a.n_sectors;
  unsigned long long* sums_s = reinterpret_cast<unsigned long long*>(smem_raw + ShootSmem<H, W>::bins_off);
  int32_t* tab_s = reinterpret_cast<int32_t*>(sums_s + n_sectors);
  int* cnts_s = tab_s + 2 * n_sectors;
  float* red_s = reinterpret_cast<float*>(smem_raw + ShootSmem<H, W>::red_off);
  const int tid = threadIdx.x;
  const int c = tid % W, br = tid / W;
  const int S = a.num_steps;
  const float dt = a.T / (float)S, mdt = -dt;
  const FluidParams fp{a.alpha, a.beta, a.gamma, 1.0f / (float)N};
  const int64_t P = prm.P;
  const int cl = max(c - 1, 0), cr = min(c + 1, W - 1);
  const float sc = (c == 0 || c == W - 1) ? 1.f : 0.5f;

  FS::init_luts(twH, twW, csH, csW, tid, NT);
  float* scr_u = prm.scratch + (size_t)blockIdx.x * 2 * prm.field;
  float* scr_m = scr_u + prm.field;
  __syncthreads();

  // work items are drawn from the ticket counter (see above); without the dynamic schedule (few pairs) CTA j takes
  // pairs j, j + G, ...
  const int64_t G = gridDim.x, cta = blockIdx.x;
  const int64_t K = prm.balanced ? G : 0;                       // tail pairs, handed over between chunks of steps
  const int n_chunks = prm.balanced ? S / kChunkSteps : 1;      // the last chunk takes the remainder of an odd S
  float* hand = prm.scratch + (size_t)G * 2 * prm.field;        // hand-over [u_s | m0] of tail pair i
  unsigned long long* ticket = reinterpret_cast<unsigned long long*>(hand + (size_t)G * 2 * prm.field);   // zeroed per launch
  int* flags = reinterpret_cast<int*>(ticket + 2);              // flags[i * n_chunks + c] = chunk c of tail pair i done
  __shared__ long long item_s;

  for (int64_t it = 0;; ++it) {
    int64_t p;
    int s0 = 0, s1 = S;
    float *hand_out = nullptr, *hand_in = nullptr;     // this item ends / starts at a hand-over
    int* flag_out = nullptr;
    if (!prm.balanced) {
      p = cta + it * G;
      if (p >= P) break;
    } else {
      if (tid == 0) item_s = (long long)atomicAdd(ticket, 1ull);
      __syncthreads();
      const int64_t tk = item_s;
      __syncthreads();                                   // everyone has read the ticket before the next draw
      if (tk < P - K) p = tk;
      else {
        const int64_t q = tk - (P - K);
        if (q >= K * n_chunks) break;
        const int ch = (int)(q / K);
        const int64_t i = q % K;
        p = P - K + i;
        s0 = ch * kChunkSteps;
        s1 = (ch == n_chunks - 1) ? S : s0 + kChunkSteps;
        float* hb = hand + (size_t)i * 2 * prm.field;
        if (s0 > 0) {
          hand_in = hb;
          if (tid == 0) {                                // the chunk before this one (drawn K tickets ago) is done
            while (*reinterpret_cast<volatile int*>(flags + i * n_chunks + ch - 1) == 0) __nanosleep(200);
            __threadfence();
          }
          __syncthreads();
        }
        if (s1 < S) { hand_out = hb; flag_out = flags + i * n_chunks + ch; }
      }
    }
    const int64_t b = p / a.T1;
    const int t = (int)(p % a.T1);
    const float* m0g;
    float* uout = a.u + (size_t)p * prm.field;
    const float* ucur = nullptr;   // u_s in global memory (nullptr == identically zero)

    if (s0 == 0) {
    // ---- m0 = flat(v0)   (or m0 given directly: lagomorph.expmap(metric, m0))
    {
      const float* f0 = a.v0 + (size_t)p * prm.field;
#pragma unroll 4
      for (int k = 0; k < NB; ++k) {
        const int r = k * RB + br;
        // v0 is read once (streaming load); a momentum input is re-read by the gathers of every step
        z[r * LD + c] = a.v0_is_momentum ? make_float2(__ldg(f0 + r * W + c), __ldg(f0 + N + r * W + c))
                                         : make_float2(B2_ONCE_LOAD(f0 + r * W + c), B2_ONCE_LOAD(f0 + N + r * W + c));
      }
    }
    __syncthreads();
    if (a.v0_is_momentum) {
      m0g = a.v0 + (size_t)p * prm.field;
    } else {
      // a pair that will be finished by another CTA keeps its momentum where that CTA finds it
      float* m0w = a.m0 ? a.m0 + (size_t)p * prm.field : (hand_out ? hand_out + prm.field : scr_m);
      fluid_smem<H, W, false, NT>(z, twH, twW, csH, csW, fp, tid);
#pragma unroll 4
      for (int k = 0; k < NB; ++k) {
        const int r = k * RB + br;
        const float2 v = z[r * LD + c];
        m0w[r * W + c] = v.x;
        m0w[N + r * W + c] = v.y;
      }
      m0g = m0w;
      __syncthreads();   // every thread has read z before the next transform overwrites it
    }
    } else {
      // ---- take over a geodesic at step s0 (its earlier chunks are done, see the wait above): reload u_{s0} into
      // the shared-memory field; m0 and u_{s0} are where the previous chunk left them
      m0g = a.v0_is_momentum ? a.v0 + (size_t)p * prm.field : (a.m0 ? a.m0 + (size_t)p * prm.field : hand_in + prm.field);
      ucur = a.traj ? a.traj + ((size_t)(s0 * 2) * P + p) * prm.field : hand_in;
#pragma unroll 4
      for (int k = 0; k < NB; ++k) {
        const int r = k * RB + br;
        z[r * LD + c] = make_float2(__ldcg(ucur + r * W + c), __ldcg(ucur + N + r * W + c));
      }
      __syncthreads();
    }

    for (int s = s0; s < s1; ++s) {
      // ---- m = Ad*_{u_s} m0 -> z.  z holds u_s (float2 per pixel, written by the previous compose):
      // the (I + Du)^T stencil reads shared memory; only the 4-tap gather of m0 goes to L1/L2.
      // Band k is overwritten with m only after every thread has read it (one barrier per band); the
      // row above the next band is prefetched into a register before it is overwritten.
#ifndef B2_EPI_PREFETCH
#define B2_EPI_PREFETCH 1     // 3.231 -> 3.207 ms at configs[1] (both options: 3.160)
#endif
#if B2_EPI_PREFETCH
      // the epilogue touches the target-frame mask and the source image for the first time: pull them into L2 one
      // EPDiff step ahead (one 128-byte line per thread)
      if (s == S - 1 && !hand_out) {
        if (a.S || LOSS) {
          const float* tp = a.tar_slice_stride ? a.tar + (size_t)b * a.tar_slice_stride + (size_t)t * N : a.tar + (size_t)p * N;
          for (int i = tid * 32; i < N; i += NT * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(tp + i));
        }
        if (a.sdef || LOSS) {
          const float* sp = a.src_per_pair
                                ? (a.src_slice_stride ? a.src + (size_t)b * a.src_slice_stride + (size_t)t * N : a.src + (size_t)p * N)
                                : a.src + (size_t)b * (a.src_slice_stride ? a.src_slice_stride : N);
          for (int i = tid * 32; i < N; i += NT * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + i));
        }
      }
#endif
#ifndef B2_ADSTAR_ROWS
#define B2_ADSTAR_ROWS 1      // 3.231 -> 3.187 ms at configs[1]
#endif
#if B2_ADSTAR_ROWS
      if (s > 0) {
        // Row-sliding form of the Ad* phase: a thread owns NB CONSECUTIVE rows of its column, so u_s(r-1), u_s(r),
        // u_s(r+1) slide through registers - three shared-memory reads per pixel (below, left, right) instead of five.
        // The whole image is read before anything is overwritten (one barrier).  Same arithmetic: bit-identical.
        const int rb0 = br * NB;
        float2 m[NB];
        float2 up = z[max(rb0 - 1, 0) * LD + c], ce = z[rb0 * LD + c];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int r = rb0 + j;
          const float2 dn = (r == H - 1) ? ce : z[(r + 1) * LD + c];
          const float2 lf = z[r * LD + cl];
          const float2 rt = z[r * LD + cr];
          const float sr = (r == 0 || r == H - 1) ? 1.f : 0.5f;
          const float d00 = sr * (dn.x - up.x), d10 = sr * (dn.y - up.y);
          const float d01 = sc * (rt.x - lf.x), d11 = sc * (rt.y - lf.y);
          float w0, w1;
          gather2<BG, kFastGather>(m0g, N, (float)r + ce.x, (float)c + ce.y, H, W, w0, w1);
          m[j] = make_float2(w0 + (d00 * w0 + d10 * w1), w1 + (d01 * w0 + d11 * w1));
          up = ce;
          ce = dn;
        }
        __syncthreads();                         // every thread has read the whole image
#pragma unroll
        for (int j = 0; j < NB; ++j) z[(rb0 + j) * LD + c] = m[j];
        __syncthreads();
      }
#else
      if (s > 0) {
#ifndef B2_ADSTAR_G
#define B2_ADSTAR_G 16
#endif
        constexpr int G = NB < B2_ADSTAR_G ? NB : B2_ADSTAR_G;   // bands per barrier: G pixels of m stay in registers
        float2 up_saved = make_float2(0.f, 0.f);
        for (int g = 0; g < NB / G; ++g) {
          float2 m[G];
          float2 up_next = up_saved;
#pragma unroll
          for (int j = 0; j < G; ++j) {
            const int k = g * G + j, r = k * RB + br;
            const float2 ce = z[r * LD + c];
            const float2 up = (br == 0) ? (j == 0 ? (g == 0 ? ce : up_saved) : z[(r - 1) * LD + c]) : z[(r - 1) * LD + c];
            const float2 dn = (r == H - 1) ? ce : z[(r + 1) * LD + c];
            const float2 lf = z[r * LD + cl];
            const float2 rt = z[r * LD + cr];
            if (j == G - 1 && br == 0 && g + 1 < NB / G) up_next = z[((k + 1) * RB - 1) * LD + c];
            const float sr = (r == 0 || r == H - 1) ? 1.f : 0.5f;
            const float d00 = sr * (dn.x - up.x), d10 = sr * (dn.y - up.y);
            const float d01 = sc * (rt.x - lf.x), d11 = sc * (rt.y - lf.y);
            float w0, w1;
            gather2<BG, kFastGather>(m0g, N, (float)r + ce.x, (float)c + ce.y, H, W, w0, w1);
            m[j] = make_float2(w0 + (d00 * w0 + d10 * w1), w1 + (d01 * w0 + d11 * w1));
          }
          __syncthreads();                       // every thread has read the rows of this group
#pragma unroll
          for (int j = 0; j < G; ++j) z[((g * G + j) * RB + br) * LD + c] = m[j];
          up_saved = up_next;
        }
        __syncthreads();
      }
#endif
      // ---- v = sharp(m)  (in shared memory)
      fluid_smem<H, W, true, NT>(z, twH, twW, csH, csW, fp, tid);

      // ---- u_{s+1} = interp(u_s, v, -dt) - dt v  -> global (gathers of the next step, outputs) and,
      // in place of v, into z (stencil of the next Ad*); trajectory / velocity outputs
      float* unext;
      if (a.traj) unext = (s + 1 < S) ? a.traj + ((size_t)((s + 1) * 2 + 0) * P + p) * prm.field : uout;
      else if (hand_out && s + 1 == s1) unext = hand_out;        // last step of a head piece: u_h for the finisher
      else unext = (((S - (s + 1)) & 1) == 0) ? uout : scr_u;
      float* vtraj = a.traj ? a.traj + ((size_t)(s * 2 + 1) * P + p) * prm.field : nullptr;
      if (s == 0) {
        float* utraj0 = a.traj ? a.traj + ((size_t)p) * prm.field : nullptr;
        float* velout = a.vel ? a.vel + (size_t)p * prm.field : nullptr;
        float acc_vm = 0.f;
#pragma unroll 2
        for (int k = 0; k < NB; ++k) {
          const int r = k * RB + br, i = r * W + c;
          const float2 v = z[r * LD + c];
          const float2 n = make_float2(mdt * v.x, mdt * v.y);
          unext[i] = n.x;
          unext[N + i] = n.y;
          z[r * LD + c] = n;
          if (utraj0) { utraj0[i] = 0.f; utraj0[N + i] = 0.f; }
          if (velout) { B2_TRAJ_STORE(velout + i, v.x); B2_TRAJ_STORE(velout + N + i, v.y); }
          if (vtraj) { B2_TRAJ_STORE(vtraj + i, v.x); B2_TRAJ_STORE(vtraj + N + i, v.y); }
          if (LOSS) acc_vm += v.x * m0g[i] + v.y * m0g[N + i];
        }
        if (LOSS) {
          // loss epilogue, regularisation term sum vel . m0: reduced and written here so that no accumulator
          // lives across the geodesic (64-register budget of the 1024-thread CTA)
          float none = 0.f;
          block_reduce2<NT>(acc_vm, none, red_s, tid);
          if (tid == 0) a.loss_terms[2 * p + 1] = acc_vm;
        }
      } else {
#pragma unroll (kComposeUnroll)
        for (int k = 0; k < NB; ++k) {
          const int r = k * RB + br, i = r * W + c;
          const float2 v = z[r * LD + c];
          float2 n;
          gather2<BG, kFastGather>(ucur, N, (float)r + mdt * v.x, (float)c + mdt * v.y, H, W, n.x, n.y);
          n.x += mdt * v.x;
          n.y += mdt * v.y;
          unext[i] = n.x;
          unext[N + i] = n.y;
          z[r * LD + c] = n;
          if (vtraj) { B2_TRAJ_STORE(vtraj + i, v.x); B2_TRAJ_STORE(vtraj + N + i, v.y); }
        }
      }
      ucur = unext;
      __syncthreads();   // u_{s+1} visible to the whole CTA (global and shared copies)
    }
    if (hand_out) {
      // chunk done: publish (u_s1, m0) to whoever draws the next chunk of this pair (release: barrier above, fence)
      if (tid == 0) {
        __threadfence();
        atomicExch(flag_out, 1);
      }
      continue;
    }

    // ---- deformed_source = interp(src, u^S)
    if (a.sdef || LOSS) {
      // Lagrangian split: one source image per slice; Eulerian split (src_per_pair): one per pair, either
      // dense (P,1,H,W) or frames of a strided cine volume (slice stride given)
      const float* src = a.src_per_pair
                             ? (a.src_slice_stride ? a.src + (size_t)b * a.src_slice_stride + (size_t)t * N
                                                   : a.src + (size_t)p * N)
                             : a.src + (size_t)b * (a.src_slice_stride ? a.src_slice_stride : N);
      float* sd = a.sdef ? a.sdef + (size_t)p * N : nullptr;
      const float* tarp = nullptr;
      if (LOSS)
        tarp = a.tar_slice_stride ? a.tar + (size_t)b * a.tar_slice_stride + (size_t)t * N : a.tar + (size_t)p * N;
      float acc_sq = 0.f;
#pragma unroll 2
      for (int k = 0; k < NB; ++k) {
        const int r = k * RB + br;
        const float2 u = z[r * LD + c];
        const float val = gather1_ldg<BG>(src, (float)r + u.x, (float)c + u.y, H, W);
        if (sd) B2_TRAJ_STORE(sd + r * W + c, val);
        if (LOSS) {
          const float d = __ldg(tarp + r * W + c) - val;
          acc_sq += d * d;
        }
      }
      if (LOSS) {
        float none = 0.f;
        block_reduce2<NT>(acc_sq, none, red_s, tid);
        if (tid == 0) a.loss_terms[2 * p] = acc_sq;
      }
    }

    // ---- strain matrix column t of slice b
    if (a.S) {
      const SectorFrame sf = sector_frame_of(a.table, a.table_slice_stride, a.theta0, a.clockwise, b);
      for (int i = tid; i < 2 * n_sectors; i += NT) tab_s[i] = sf.table[i];      // this slice's (rotated) boundaries
      for (int i = tid; i < n_sectors; i += NT) { sums_s[i] = 0ull; cnts_s[i] = 0; }
#ifndef B2_STRAIN_COMPACT
#define B2_STRAIN_COMPACT 1
#endif
      int* n_members = reinterpret_cast<int*>(red_s);         // free here: the loss reductions are done
      if (tid == 0) *n_members = 0;
      __syncthreads();                                        // also: every thread is done with z (warp phase)
      const float* tarp = a.tar_slice_stride ? a.tar + (size_t)b * a.tar_slice_stride + (size_t)t * N
                                             : a.tar + (size_t)p * N;
#if B2_STRAIN_COMPACT
      // z is dead from here to the next pair: it holds the list of member pixels
      strain_bin_frame_compact<NT>(ucur, ucur + N, tarp, reinterpret_cast<const long long*>(a.moments) + 3 * b,
                                   tab_s, n_sectors, H, W, sums_s, cnts_s, tid, sf.theta0, sf.flip,
                                   reinterpret_cast<unsigned short*>(z), n_members);
#else
      strain_bin_frame<NT>(ucur, ucur + N, tarp, reinterpret_cast<const long long*>(a.moments) + 3 * b,
                           tab_s, n_sectors, H, W, sums_s, cnts_s, tid, sf.theta0, sf.flip);
#endif
      strain_store_column<NT>(sums_s, cnts_s, a.S, a.counts, (int)b, t, (int)a.T1, n_sectors, a.n_frames, tid);
    }
    __syncthreads();   // z and bins free for the next pair
  }
}

// Resident CTAs per SM of the fused kernel as the runtime reports it (shared memory, registers and threads all
// count): the persistent grid is #SMs x this, and the per-CTA scratch is sized from the same number.
template <int H, int W, int NT>
struct FusedCfg {
  static int ctas_per_sm() {
    static int cached = 0;
    if (cached) return cached;
    const size_t smem = ShootSmem<H, W>::bytes;
    int per = 0;
    if (cudaFuncSetAttribute(shoot_fwd_kernel<H, W, NT, B2_BG_CLAMP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, shoot_fwd_kernel<H, W, NT, B2_BG_CLAMP, false>, NT, smem) !=
            cudaSuccess ||
        per < 1) {
      (void)cudaGetLastError();   // no device / query failed: fall back to the shared-memory bound
      per = (int)((224 * 1024) / (smem + 1024));
      if (per < 1) per = 1;
      if (per * NT > 2048) per = 2048 / NT;
      return per;                 // not cached: a later call with a device may do better
    }
    cached = per;
    return per;
  }
};

static int sm_count() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return sms;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 148;
  return sms;
}

static bool fused_size(int64_t H, int64_t W, int flags) {
  return H == W && (H == 16 || H == 32 || H == 64 || H == 128) && !(flags & B2_FLAG_OPLEVEL);
}

static int64_t fused_grid(int64_t P, int64_t H) {
  int per = 1;
  switch ((int)H) {
    case 16: per = FusedCfg<16, 16, 128>::ctas_per_sm(); break;
    case 32: per = FusedCfg<32, 32, 256>::ctas_per_sm(); break;
    case 64: per = FusedCfg<64, 64, 256>::ctas_per_sm(); break;
    case 128: per = FusedCfg<128, 128, 1024>::ctas_per_sm(); break;   // 1024 threads measured 22 % faster than 512
  }
  int64_t g = (int64_t)sm_count() * per;
  return g < P ? g : P;
}

template <int H, int W, int NT, int BG, bool LOSS>
static int launch_fused_variant(const ShootParams& prm, int64_t grid, cudaStream_t st) {
  const size_t smem = ShootSmem<H, W>::bytes;
  B2_CUDA(cudaFuncSetAttribute(shoot_fwd_kernel<H, W, NT, BG, LOSS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  shoot_fwd_kernel<H, W, NT, BG, LOSS><<<(unsigned)grid, NT, smem, st>>>(prm);
  B2_CHECK_LAUNCH();
  return B2_OK;
}

// the loss epilogue is a compile-time variant: the inference kernel carries none of its branches
template <int H, int W, int NT>
static int launch_fused(const ShootParams& prm, int64_t grid, cudaStream_t st) {
  const bool loss = prm.a.loss_terms != nullptr;
  if (prm.a.background == B2_BG_CLAMP)
    return loss ? launch_fused_variant<H, W, NT, B2_BG_CLAMP, true>(prm, grid, st)
                : launch_fused_variant<H, W, NT, B2_BG_CLAMP, false>(prm, grid, st);
  return loss ? launch_fused_variant<H, W, NT, B2_BG_ZERO, true>(prm, grid, st)
              : launch_fused_variant<H, W, NT, B2_BG_ZERO, false>(prm, grid, st);
}

#ifndef B2_BWD_U1
#define B2_BWD_U1 4
#endif
#ifndef B2_BWD_U3A
#define B2_BWD_U3A 8
#endif
#ifndef B2_BWD_U3B
#define B2_BWD_U3B 4
#endif
constexpr int kBwdUnrollB1 = B2_BWD_U1, kBwdUnrollB3a = B2_BWD_U3A, kBwdUnrollB3b = B2_BWD_U3B;
#ifndef B2_BWD_PREFETCH
#define B2_BWD_PREFETCH 0
#endif
#ifndef B2_BWD_PIPE
#define B2_BWD_PIPE 2     // software-pipeline depth (rows) of the compose adjoint's trajectory loads; 0 = off
                          // (configs[2] training step: off 11.84 ms, 2: 11.61, 4: 11.66, 8: 11.62)
#endif

#ifndef B2_BWD_EVICT_FIRST
#define B2_BWD_EVICT_FIRST 0
#endif
#ifndef B2_BWD_ROWWIN
#define B2_BWD_ROWWIN 1   // sliding register window over the rows of a thread's column in the Ad* adjoint passes
#endif
// Trajectory loads of the adjoint with an L2 evict-first policy (createpolicy + ld.global.nc.L2::cache_hint): the
// (u_s, v_s) entry streams through once per step, the scratch fields and m0 are what should stay in L2.
__device__ __forceinline__ unsigned long long make_evict_first_policy() {
  unsigned long long pol = 0;
#if B2_BWD_EVICT_FIRST
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#endif
  return pol;
}
#ifndef B2_BWD_VS_STREAM
#define B2_BWD_VS_STREAM 0     // 1: ld.global.cs for v_s - measured neutral (10.15 vs 10.15 ms per training step)
#endif
// v_s is read exactly once per adjoint step (u_s is gathered: its lines are re-used)
__device__ __forceinline__ float ldg_once(const float* p) {
#if B2_BWD_VS_STREAM
  return __ldcs(p);
#else
  return __ldg(p);
#endif
}
__device__ __forceinline__ float ldg_stream(const float* p, unsigned long long pol) {
#if B2_BWD_EVICT_FIRST
  float v;
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
#else
  (void)pol;
  return __ldg(p);
#endif
}

// L2 prefetch of one (u_s, v_s) trajectory entry (2 fields): the adjoint's first phase of a step touches them for
// the first time (HBM latency on a dependent load chain); issued one phase earlier they arrive in L2 in time.
// Variant 1: one prefetch.global.L2 per 128-byte line, spread over the CTA; variant 2: four bulk prefetches.
template <int NT>
__device__ __forceinline__ void prefetch_traj_l2(const float* us, const float* vs, int field_floats, int tid) {
#if B2_BWD_PREFETCH == 1
  for (int i = tid * 32; i < field_floats; i += NT * 32) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(us + i));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(vs + i));
  }
#elif B2_BWD_PREFETCH == 2
  if (tid == 0) {
    const unsigned bytes = (unsigned)(field_floats * sizeof(float));
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(us), "r"(bytes));
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(vs), "r"(bytes));
  }
#else
  (void)us; (void)vs; (void)field_floats; (void)tid;
#endif
}

// shared memory of the adjoint kernel: field + LUTs, then the seed prologue's sector table and weights
template <int H, int W>
struct BwdSmem {
  static constexpr size_t seed_off = (FluidSmem<H, W>::bytes + 15) & ~size_t(15);
  static constexpr size_t bytes = seed_off + sizeof(int32_t) * 3 * kFusedMaxSectors;
};

// ------------------------------------------------------------------ fused EPDiff adjoint (path A)
// Reverse sweep over the saved trajectory, one CTA per frame-pair at a time, same residency scheme as the
// forward: dL/dv_s -> dL/dm_s lives in shared memory (the self-adjoint sharp is the in-SM FFT), the gradient
// accumulators dL/du (ping-pong), dL/dm0 and w = m0 o (id + u_s) sit in per-CTA global scratch that stays in L2.
// Buffers that receive float atomics (RED goes to L2) are only ever READ with ld.global.cg, so a stale L1 line
// can never be observed; u_s, v_s and m0 are read-only for this kernel (ld.global.nc).

template <int H, int W, int NT, int BG>
__global__ void __launch_bounds__(NT)
shoot_bwd_kernel(const ShootBwdParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using FS = FluidSmem<H, W>;
  static_assert(NT % W == 0 && H % (NT / W) == 0, "CTA must tile the grid in whole row bands");
  constexpr int LD = FS::LD, N = H * W, RB = NT / W, NB = H / RB;
  float2 *z, *twH, *twW, *csH, *csW;
  FS::carve(smem_raw, z, twH, twW, csH, csW);
  const int tid = threadIdx.x, lane = tid & 31;
  // Thread <-> pixel map of the adjoint: a thread keeps one column c and walks the NB CONSECUTIVE rows
  // rbase + k (the forward kernel strides rows by RB instead), lanes along the contiguous axis.  The 2x2 splat
  // footprints of vertically adjacent pixels then overlap inside one thread and those of horizontally adjacent
  // pixels inside one warp: splat2_agg merges them before anything goes to the L2 atomic unit.
  const int c = tid % W, rbase = (tid / W) * NB;
  const int S = prm.num_steps;
  const float mdt = -prm.T / (float)S;
  const FluidParams fp{prm.alpha, prm.beta, prm.gamma, 1.0f / (float)N};
  const int64_t P = prm.P;
  const int cl = max(c - 1, 0), cr = min(c + 1, W - 1);
  const float sc = diff_scale(c, W);
  const float cmc = (c >= 1) ? diff_scale(c - 1, W) : 0.f, cpc = (c <= W - 2) ? diff_scale(c + 1, W) : 0.f;
  const float c0c = (c == W - 1 ? 1.f : 0.f) - (c == 0 ? 1.f : 0.f);
  FS::init_luts(twH, twW, csH, csW, tid, NT);
  const unsigned long long pol = make_evict_first_policy();
  const BwdSeed& sd = prm.seed;
  int32_t* tab_s = reinterpret_cast<int32_t*>(smem_raw + BwdSmem<H, W>::seed_off);      // seeds: sector table
  float* gk_s = reinterpret_cast<float*>(tab_s + 2 * kFusedMaxSectors);                   //        + per-sector weights
  // Work items come from the same kind of ticket counter as in shoot_fwd_kernel: whole pairs first, then the last
  // K = grid-size pairs in chunks of `chunk_steps` adjoint steps, so that no CTA is left with a whole pair while the
  // others are done.  The state of the reverse sweep at a step boundary is (dL/du_s, dL/dm0), all of it in the scratch
  // and written through L2 (stcg / RED): a tail pair uses three scratch fields of its own (slots G .. 2G-1, behind the
  // per-CTA ones), so a chunk continues where the previous one stopped - whoever ran it.  Thread 0 draws and decodes
  // the item; the descriptor lives in shared memory and is read where it is needed (the kernel sits at its register
  // limit: a descriptor in registers spilled inside the hot loops and cost more than the schedule gained).
  __shared__ int item_sh[6];      // pair, first step, last step, scratch slot, flag to publish (-1: none), round
  if (tid == 0) item_sh[5] = 0;
  __syncthreads();

  for (;;) {
    if (tid == 0) {
      const int64_t G = gridDim.x;
      int64_t pp = -1;
      int s_hi = S - 1, s_lo = 0, pub = -1;
      int64_t slot = blockIdx.x;
      if (!prm.ticket) {
        const int64_t q = blockIdx.x + (int64_t)item_sh[5] * G;
        if (q < P) pp = q;
        item_sh[5] += 1;
      } else {
        const int cs = prm.chunk_steps, n_chunks = S / cs;       // the last chunk takes the remainder
        const int64_t tk = (int64_t)atomicAdd(prm.ticket, 1ull);
        if (tk < P - G) pp = tk;
        else if (tk - (P - G) < G * n_chunks) {
          const int64_t q = tk - (P - G);
          const int ch = (int)(q / G);
          const int64_t i = q % G;
          pp = P - G + i;
          slot = G + i;
          s_hi = S - 1 - ch * cs;
          s_lo = (ch == n_chunks - 1) ? 0 : s_hi - cs + 1;
          if (ch > 0) {                                  // the chunk before this one (drawn G tickets ago)
            while (*reinterpret_cast<volatile int*>(prm.flags + i * n_chunks + ch - 1) == 0) __nanosleep(200);
            __threadfence();                             // acquire
          }
          if (s_lo > 0) pub = (int)(i * n_chunks + ch);
        }
      }
      item_sh[0] = (int)pp; item_sh[1] = s_hi; item_sh[2] = s_lo; item_sh[3] = (int)slot; item_sh[4] = pub;
    }
    __syncthreads();
    if (item_sh[0] < 0) break;
    const int64_t p = item_sh[0];
    const int s_hi = item_sh[1];
    float* Ga = prm.scratch + (size_t)item_sh[3] * 3 * prm.field;
    float* Gb = Ga + prm.field;
    float* A = Gb + prm.field;
    const float* m0p = prm.m0 + (size_t)p * prm.field;
    // every adjoint step with s > 0 swaps the two dL/du buffers: after the steps S-1 .. s_hi+1 the current one is
    const bool swapped = ((S - 1 - s_hi) & 1) != 0;
    float* Gcur = swapped ? Gb : Ga;
    float* Gnext = swapped ? Ga : Gb;
    if (s_hi == S - 1) {                              // first chunk of the pair: seed dL/du^S and dL/dm0
    for (int k = 0; k < NB; ++k) {
      const int i = (rbase + k) * W + c;
      Gcur[i] = prm.gu ? __ldg(prm.gu + (size_t)p * prm.field + i) : 0.f;
      Gcur[N + i] = prm.gu ? __ldg(prm.gu + (size_t)p * prm.field + N + i) : 0.f;
      A[i] = prm.gm0 ? __ldg(prm.gm0 + (size_t)p * prm.field + i) : 0.f;
      A[N + i] = prm.gm0 ? __ldg(prm.gm0 + (size_t)p * prm.field + N + i) : 0.f;
    }
    if (sd.gS || sd.g_sq) {
      // ---- seeds of dL/du^S taken here instead of by two kernels through a gradient image: the adjoint of the
      // strain-matrix reduction (REDs of the member pixels' stencils into the accumulator) and of the squared
      // error of the warped source (own pixel, recomputed from the taps of src at x + u^S)
      const int64_t b = p / sd.T1;
      const int t = (int)(p % sd.T1);
      const float* uS = sd.uS + (size_t)p * prm.field;
      const float* tarp = sd.tar_slice_stride ? sd.tar + (size_t)b * sd.tar_slice_stride + (size_t)t * N
                                              : sd.tar + (size_t)p * N;
      if (sd.gS) {
        const SectorFrame sf = sector_frame_of(sd.table, sd.table_slice_stride, sd.theta0, sd.clockwise, b);
        for (int i = tid; i < 2 * sd.n_sectors; i += NT) tab_s[i] = sf.table[i];
        strain_bwd_weights<NT>(sd.gS, sd.counts, (int)b, t, (int)sd.T1, sd.n_sectors, sd.n_frames, gk_s, tid);
        __syncthreads();            // accumulator seeded, table and weights visible
#ifndef B2_BWD_SEED_COMPACT
#define B2_BWD_SEED_COMPACT 1
#endif
        // the field buffer z is not in use before the first compose adjoint: it holds the list of member pixels
        strain_bwd_frame<NT>(uS, uS + N, tarp, sd.moments + 3 * b, tab_s, sd.n_sectors, H, W, gk_s, Gcur, Gcur + N, tid,
                             sf.theta0, sf.flip, B2_BWD_SEED_COMPACT ? reinterpret_cast<unsigned short*>(z) : nullptr);
      }
      if (sd.g_sq) {
        const float* srcp = sd.src_per_pair
                                ? (sd.src_slice_stride ? sd.src + (size_t)b * sd.src_slice_stride + (size_t)t * N
                                                       : sd.src + (size_t)p * N)
                                : sd.src + (size_t)b * (sd.src_slice_stride ? sd.src_slice_stride : N);
        const float g2 = 2.f * __ldg(sd.g_sq + p);
        for (int k = 0; k < NB; ++k) {
          const int r = rbase + k, i = r * W + c;
          const Taps tp = make_taps<BG>((float)r + __ldg(uS + i), (float)c + __ldg(uS + N + i), H, W);
          const float v00 = __ldg(srcp + tp.o00), v10 = __ldg(srcp + tp.o10), v01 = __ldg(srcp + tp.o01), v11 = __ldg(srcp + tp.o11);
          const float g = g2 * (tap_sample<BG>(tp, v00, v10, v01, v11) - __ldg(tarp + i));
          float a0, a1;
          tap_grad<BG>(tp, v00, v10, v01, v11, a0, a1);
          red_add(Gcur + i, g * a0);       // RED: the strain adjoint of neighbouring pixels may target this pixel too
          red_add(Gcur + N + i, g * a1);
        }
      }
    }
    }
    __syncthreads();

    for (int s = s_hi; s >= *reinterpret_cast<volatile int*>(item_sh + 2); --s) {
      const float* us = prm.traj + ((size_t)(2 * s) * P + p) * prm.field;
      const float* vs = prm.traj + ((size_t)(2 * s + 1) * P + p) * prm.field;
      if (s > 0) {
        // ---- adjoint of u_{s+1} = interp(u_s, v_s, -dt) - dt v_s : dL/dv_s -> z, splat of dL/du_{s+1} -> Gnext
        for (int k = 0; k < NB; ++k) {
          const int i = (rbase + k) * W + c;
          Gnext[i] = 0.f;
          Gnext[N + i] = 0.f;
        }
        __syncthreads();
        SplatCarry cy{-1, 0.f, 0.f};
#if B2_BWD_PIPE
        // (u_s, v_s) are touched here for the first time: HBM latency on a dependent chain v_s -> taps -> u_s.
        // Software pipeline over groups of kPipe rows: the v_s values of the NEXT group are loaded and the u_s
        // lines around its pixels (the taps are within a pixel of them for |dt v| < 1) are pulled into L1 while
        // the current group is processed.
        constexpr int kPipe = B2_BWD_PIPE < NB ? B2_BWD_PIPE : NB;
        static_assert(NB % kPipe == 0, "pipeline depth must divide the rows per thread");
        float va[kPipe], vb[kPipe];
#pragma unroll
        for (int j = 0; j < kPipe; ++j) {
          const int i = (rbase + j) * W + c;
          va[j] = ldg_once(vs + i);
          vb[j] = ldg_once(vs + N + i);
          asm volatile("prefetch.global.L1 [%0];" ::"l"(us + i));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(us + N + i));
        }
        for (int k0 = 0; k0 < NB; k0 += kPipe) {
          float na[kPipe], nb[kPipe];
#pragma unroll
          for (int j = 0; j < kPipe; ++j) {
            const int kn = min(k0 + kPipe + j, NB - 1);          // last group: harmless re-read of its own rows
            const int i = (rbase + kn) * W + c;
            na[j] = ldg_once(vs + i);
            nb[j] = ldg_once(vs + N + i);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(us + i));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(us + N + i));
          }
#pragma unroll
          for (int j = 0; j < kPipe; ++j) {
            const int r = rbase + k0 + j, i = r * W + c;
            const float g0 = __ldcg(Gcur + i), g1 = __ldcg(Gcur + N + i);
            const float v0 = va[j], v1 = vb[j];
            const Taps t = make_taps<BG>((float)r + mdt * v0, (float)c + mdt * v1, H, W);
            float a0, a1, b0, b1;
            tap_grad<BG>(t, ldg_stream(us + t.o00, pol), ldg_stream(us + t.o10, pol), ldg_stream(us + t.o01, pol), ldg_stream(us + t.o11, pol), a0, a1);
            tap_grad<BG>(t, ldg_stream(us + N + t.o00, pol), ldg_stream(us + N + t.o10, pol), ldg_stream(us + N + t.o01, pol), ldg_stream(us + N + t.o11, pol), b0, b1);
            z[r * LD + c] = make_float2(mdt * (g0 * a0 + g1 * b0 + g0), mdt * (g0 * a1 + g1 * b1 + g1));
            splat2_agg<BG>(Gnext, N, t, g0, g1, cy, lane);
          }
#pragma unroll
          for (int j = 0; j < kPipe; ++j) { va[j] = na[j]; vb[j] = nb[j]; }
        }
#else
#pragma unroll (kBwdUnrollB1)
        for (int k = 0; k < NB; ++k) {
          const int r = rbase + k, i = r * W + c;
          const float g0 = __ldcg(Gcur + i), g1 = __ldcg(Gcur + N + i);
          const float v0 = ldg_stream(vs + i, pol), v1 = ldg_stream(vs + N + i, pol);
          const Taps t = make_taps<BG>((float)r + mdt * v0, (float)c + mdt * v1, H, W);
          float a0, a1, b0, b1;
          tap_grad<BG>(t, ldg_stream(us + t.o00, pol), ldg_stream(us + t.o10, pol), ldg_stream(us + t.o01, pol), ldg_stream(us + t.o11, pol), a0, a1);
          tap_grad<BG>(t, ldg_stream(us + N + t.o00, pol), ldg_stream(us + N + t.o10, pol), ldg_stream(us + N + t.o01, pol), ldg_stream(us + N + t.o11, pol), b0, b1);
          z[r * LD + c] = make_float2(mdt * (g0 * a0 + g1 * b0 + g0), mdt * (g0 * a1 + g1 * b1 + g1));
          splat2_agg<BG>(Gnext, N, t, g0, g1, cy, lane);
        }
#endif
        splat_flush(Gnext, N, cy);
      } else {
        // u_0 = 0: u_1 = -dt v_0, so dL/dv_0 = -dt dL/du_1 (+ the direct gradient of the velocity output)
        const float* gv = prm.gvel ? prm.gvel + (size_t)p * prm.field : nullptr;
        for (int k = 0; k < NB; ++k) {
          const int r = rbase + k, i = r * W + c;
          float a = mdt * __ldcg(Gcur + i), b = mdt * __ldcg(Gcur + N + i);
          if (gv) { a += __ldg(gv + i); b += __ldg(gv + N + i); }
          z[r * LD + c] = make_float2(a, b);
        }
      }
      __syncthreads();
      // ---- dL/dm_s = sharp(dL/dv_s)   (self-adjoint)
      fluid_smem<H, W, true, NT>(z, twH, twW, csH, csW, fp, tid);
      if (s > 0) {
        // ---- adjoint of m_s = (I + Du_s)^T (m0 o (id + u_s)), w = m0 o (id + u_s), g = dL/dm_s:
        //   dL/dm0 += splat_{x + u_s}((I + Du_s) g)                                  (REDs into A)
        //   dL/du_s += (I + Du_s) g . grad m0(x + u_s)  +  D_0^T (g_0 w) + D_1^T (g_1 w)
        // Pass A does everything that is local to a pixel with ONE gather of m0 (value and gradient from the same
        // four taps) and leaves the four products g_a w_b behind: (g_1 w_0, g_1 w_1) in place of g in shared
        // memory (column neighbours), (g_0 w_0, g_0 w_1) in the buffer of the consumed dL/du_{s+1} (row
        // neighbours; three fields of scratch per CTA).  Pass B adds the transposed differences of the products.
        float* Qb = Gcur;
        if (s >= 2)      // the next adjoint step reads (u_{s-1}, v_{s-1}) first thing
          prefetch_traj_l2<NT>(prm.traj + ((size_t)(2 * s - 2) * P + p) * prm.field,
                               prm.traj + ((size_t)(2 * s - 1) * P + p) * prm.field, (int)prm.field, tid);
        SplatCarry cy{-1, 0.f, 0.f};
#if B2_BWD_ROWWIN
        // the thread walks consecutive rows of one column: u_s(r-1), u_s(r), u_s(r+1) of that column slide through
        // registers - two loads per row (the new row, both planes) instead of six; same values, same arithmetic
        float ua_up = ldg_stream(us + max(rbase - 1, 0) * W + c, pol), ub_up = ldg_stream(us + N + max(rbase - 1, 0) * W + c, pol);
        float ua_c = ldg_stream(us + rbase * W + c, pol), ub_c = ldg_stream(us + N + rbase * W + c, pol);
#endif
#pragma unroll (kBwdUnrollB3a)
        for (int k = 0; k < NB; ++k) {
          const int r = rbase + k, i = r * W + c;
          // the compose adjoint's REDs into Gnext completed before the barriers above: the own-pixel parts of
          // dL/du_s are plain read-modify-writes (L2 path on both sides), no atomic
          const float gn0 = __ldcg(Gnext + i), gn1 = __ldcg(Gnext + N + i);
          const int ru = max(r - 1, 0), rd = min(r + 1, H - 1);
          const int oup = ru * W + c, odn = rd * W + c, olf = r * W + cl, ort = r * W + cr;
          const float sr = diff_scale(r, H);
#if B2_BWD_ROWWIN
          const float ua_dn = ldg_stream(us + odn, pol), ub_dn = ldg_stream(us + N + odn, pol);
          const float d00 = sr * (ua_dn - ua_up), d10 = sr * (ub_dn - ub_up);
          const float uc0 = ua_c, uc1 = ub_c;
          ua_up = ua_c; ub_up = ub_c; ua_c = ua_dn; ub_c = ub_dn;
          (void)oup;
#else
          const float d00 = sr * (ldg_stream(us + odn, pol) - ldg_stream(us + oup, pol)), d10 = sr * (ldg_stream(us + N + odn, pol) - ldg_stream(us + N + oup, pol));
          const float uc0 = ldg_stream(us + i, pol), uc1 = ldg_stream(us + N + i, pol);
#endif
          const float d01 = sc * (ldg_stream(us + ort, pol) - ldg_stream(us + olf, pol)), d11 = sc * (ldg_stream(us + N + ort, pol) - ldg_stream(us + N + olf, pol));
          const float2 g = z[r * LD + c];
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
          __stcg(Gnext + i, gn0 + o0);
          __stcg(Gnext + N + i, gn1 + o1);
          Qb[i] = g.x * w0;
          Qb[N + i] = g.x * w1;
          z[r * LD + c] = make_float2(g.y * w0, g.y * w1);     // own pixel only: nobody else reads g
        }
        splat_flush(A, N, cy);
        __syncthreads();
#if B2_BWD_ROWWIN
        float qa_up = Qb[max(rbase - 1, 0) * W + c], qb_up = Qb[N + max(rbase - 1, 0) * W + c];
        float qa_c = Qb[rbase * W + c], qb_c = Qb[N + rbase * W + c];
#endif
#pragma unroll (kBwdUnrollB3b)
        for (int k = 0; k < NB; ++k) {
          const int r = rbase + k, i = r * W + c;
          const float gn0 = __ldcg(Gnext + i), gn1 = __ldcg(Gnext + N + i);
          const int ru = max(r - 1, 0), rd = min(r + 1, H - 1);
          const int oup = ru * W + c, odn = rd * W + c;
          // gathered transposed differences (diffT): rows from the scratch products, columns from shared memory
          const float cmr = (r >= 1) ? diff_scale(r - 1, H) : 0.f, cpr = (r <= H - 2) ? diff_scale(r + 1, H) : 0.f;
          const float c0r = (r == H - 1 ? 1.f : 0.f) - (r == 0 ? 1.f : 0.f);
          const float2 ql = z[r * LD + cl], qr = z[r * LD + cr];
#if B2_BWD_ROWWIN
          const float qa_dn = Qb[odn], qb_dn = Qb[N + odn];      // row-neighbour products slide through registers too
          float o0 = (cmr * qa_up - cpr * qa_dn) + (cmc * ql.x - cpc * qr.x);
          float o1 = (cmr * qb_up - cpr * qb_dn) + (cmc * ql.y - cpc * qr.y);
          if (c0r != 0.f) { o0 += c0r * qa_c; o1 += c0r * qb_c; }                          // first / last image row
          qa_up = qa_c; qb_up = qb_c; qa_c = qa_dn; qb_c = qb_dn;
          (void)oup;
#else
          float o0 = (cmr * Qb[oup] - cpr * Qb[odn]) + (cmc * ql.x - cpc * qr.x);
          float o1 = (cmr * Qb[N + oup] - cpr * Qb[N + odn]) + (cmc * ql.y - cpc * qr.y);
          if (c0r != 0.f) { o0 += c0r * Qb[i]; o1 += c0r * Qb[N + i]; }                    // first / last image row
#endif
          if (c0c != 0.f) { const float2 qc = z[r * LD + c]; o0 += c0c * qc.x; o1 += c0c * qc.y; }   // first / last column
          __stcg(Gnext + i, gn0 + o0);
          __stcg(Gnext + N + i, gn1 + o1);
        }
        float* tmp = Gcur; Gcur = Gnext; Gnext = tmp;
      } else {
        // m_0 = Ad*_0 m0 = m0 exactly: dL/dm0 = accumulated splats + dL/dm_0, summed in place in shared memory
        // (the REDs into A completed before the barriers above)
        for (int k = 0; k < NB; ++k) {
          const int r = rbase + k, i = r * W + c;
          float2 g = z[r * LD + c];
          g.x += __ldcg(A + i);
          g.y += __ldcg(A + N + i);
          z[r * LD + c] = g;
        }
      }
      __syncthreads();
    }
    if (item_sh[4] >= 0) {
      // chunk done (the barrier closing its last step is behind us): publish (dL/du_s, dL/dm0) to whoever draws
      // the next chunk of this pair
      const int pub = item_sh[4];
      __syncthreads();                                   // descriptor read by everyone before thread 0 redraws
      if (tid == 0) {
        __threadfence();
        atomicExch(prm.flags + pub, 1);
      }
      continue;
    }
    // ---- dL/dv0 = flat(dL/dm0)  (or dL/dm0 itself when the forward input was the momentum)
    if (!prm.v0_is_momentum) fluid_smem<H, W, false, NT>(z, twH, twW, csH, csW, fp, tid);
    float* out = prm.gv0 + (size_t)p * prm.field;
    // d<sharp(m0), m0>/dm0 = 2 vel and flat(2 vel) = 2 m0: the regularisation gradient needs no transform
    const float g2 = prm.g_reg ? 2.f * __ldg(prm.g_reg + p) : 0.f;
    const float* radd = prm.v0_is_momentum ? prm.traj + ((size_t)P + p) * prm.field : m0p;   // v_0 of the trajectory
    for (int k = 0; k < NB; ++k) {
      const int r = rbase + k, i = r * W + c;
      float2 v = z[r * LD + c];
      if (prm.g_reg) { v.x += g2 * __ldg(radd + i); v.y += g2 * __ldg(radd + N + i); }
      B2_TRAJ_STORE(out + i, v.x);
      B2_TRAJ_STORE(out + N + i, v.y);
    }
    __syncthreads();
  }
}

// Resident CTAs per SM of the adjoint kernel as the runtime reports it; the per-CTA scratch is sized from it.
template <int H, int W, int NT>
static int bwd_ctas_per_sm() {
  static int cached = 0;
  if (cached) return cached;
  const size_t smem = BwdSmem<H, W>::bytes;
  int per = 0;
  if (cudaFuncSetAttribute(shoot_bwd_kernel<H, W, NT, B2_BG_CLAMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
          cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, shoot_bwd_kernel<H, W, NT, B2_BG_CLAMP>, NT, smem) != cudaSuccess ||
      per < 1) {
    (void)cudaGetLastError();   // no device / query failed: shared-memory bound (not cached)
    per = (int)((224 * 1024) / (smem + 1024));
    if (per < 1) per = 1;
    if (per * NT > 2048) per = 2048 / NT;
    return per;
  }
  cached = per;
  return per;
}

static int64_t fused_bwd_grid(int64_t P, int64_t H) {
  int per = 1;
  switch ((int)H) {
    case 16: per = bwd_ctas_per_sm<16, 16, 128>(); break;
    case 32: per = bwd_ctas_per_sm<32, 32, 256>(); break;
    case 64: per = bwd_ctas_per_sm<64, 64, 256>(); break;
    case 128: per = bwd_ctas_per_sm<128, 128, 1024>(); break;
  }
  const int64_t g = (int64_t)sm_count() * per;
  return g < P ? g : P;
}

#ifndef B2_BWD_L2_PERSIST
#define B2_BWD_L2_PERSIST 0
#endif
// Optional: pin the per-CTA scratch of the adjoint in L2 (persisting access-policy window on the launch stream) so
// that the streaming trajectory does not evict it.  Returns true when a window was set (reset after the launch).
static bool l2_window_set(cudaStream_t st, void* base, size_t bytes) {
#if B2_BWD_L2_PERSIST
  static int max_persist = -1, max_window = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (max_persist < 0) {
    if (cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev) != cudaSuccess) max_persist = 0;
    if (cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev) != cudaSuccess) max_window = 0;
    if (max_persist > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist) != cudaSuccess) max_persist = 0;
    (void)cudaGetLastError();
  }
  if (max_persist <= 0 || max_window <= 0) return false;
  cudaStreamAttrValue v = {};
  v.accessPolicyWindow.base_ptr = base;
  v.accessPolicyWindow.num_bytes = bytes < (size_t)max_window ? bytes : (size_t)max_window;
  const double r = (double)max_persist / (double)v.accessPolicyWindow.num_bytes;
  v.accessPolicyWindow.hitRatio = r < 1.0 ? (float)r : 1.0f;
  v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) { (void)cudaGetLastError(); return false; }
  return true;
#else
  (void)st; (void)base; (void)bytes;
  return false;
#endif
}
static void l2_window_reset(cudaStream_t st) {
  cudaStreamAttrValue v = {};
  v.accessPolicyWindow.num_bytes = 0;
  (void)cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v);
}

template <int H, int W, int NT>
static int launch_fused_bwd(const ShootBwdParams& prm_in, int background, cudaStream_t st) {
  const size_t smem = BwdSmem<H, W>::bytes;
  const int64_t grid = fused_bwd_grid(prm_in.P, H);
  ShootBwdParams prm = prm_in;
  // dynamic schedule: worth it once every CTA has at least two pairs; at most kBwdMaxChunks chunks per tail pair (the
  // workspace query does not know the step count).  Ticket + flags sit behind the 2 x grid scratch slots.
  int cs = kChunkSteps;
  while (prm.num_steps / cs > kBwdMaxChunks) ++cs;
  prm.ticket = nullptr; prm.flags = nullptr; prm.chunk_steps = cs;
  if (B2_BALANCED_BWD && prm.P >= 2 * grid && prm.num_steps >= 2 * cs) {
    unsigned char* tk = reinterpret_cast<unsigned char*>(prm.scratch) + align256(sizeof(float) * (size_t)grid * 6 * prm.field);
    prm.ticket = reinterpret_cast<unsigned long long*>(tk);
    prm.flags = reinterpret_cast<int*>(tk + 16);
    B2_CUDA(cudaMemsetAsync(tk, 0, 16 + sizeof(int) * (size_t)grid * kBwdMaxChunks, st));
  }
  const bool win = l2_window_set(st, prm.scratch, sizeof(float) * (size_t)grid * 3 * prm.field);
  if (background == B2_BG_CLAMP) {
    B2_CUDA(cudaFuncSetAttribute(shoot_bwd_kernel<H, W, NT, B2_BG_CLAMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    shoot_bwd_kernel<H, W, NT, B2_BG_CLAMP><<<(unsigned)grid, NT, smem, st>>>(prm);
  } else {
    B2_CUDA(cudaFuncSetAttribute(shoot_bwd_kernel<H, W, NT, B2_BG_ZERO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    shoot_bwd_kernel<H, W, NT, B2_BG_ZERO><<<(unsigned)grid, NT, smem, st>>>(prm);
  }
  if (win) l2_window_reset(st);
  B2_CHECK_LAUNCH();
  return B2_OK;
}

// ------------------------------------------------------------------ small elementwise helpers
// y[p, :] += 2 g[p] x[p, :]   (closed-form regularisation gradient on the op-level path)
__global__ void add_scaled_pairs_kernel(float* __restrict__ y, const float* __restrict__ x, const float* __restrict__ g,
                                        int64_t P, int field) {
  for (int64_t p = blockIdx.y; p < P; p += gridDim.y) {
    const float g2 = 2.f * __ldg(g + p);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < field; i += gridDim.x * blockDim.x)
      y[(size_t)p * field + i] += g2 * __ldg(x + (size_t)p * field + i);
  }
}
__global__ void axpby_kernel(float* __restrict__ y, const float* __restrict__ x, float a, float b, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = a * x[i] + b * y[i];
}
static int axpby(float* y, const float* x, float a, float b, size_t n, cudaStream_t st) {
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  axpby_kernel<<<(unsigned)blocks, 256, 0, st>>>(y, x, a, b, n);
  B2_CHECK_LAUNCH();
  return B2_OK;
}


}  // namespace b2

using namespace b2;

extern "C" int64_t b2_shoot_workspace_bytes(int64_t B, int64_t T1, int64_t H, int64_t W, int num_steps) {
  return b2_shoot_workspace_bytes_flags(B, T1, H, W, num_steps, 0);
}

extern "C" int64_t b2_shoot_workspace_bytes_flags(int64_t B, int64_t T1, int64_t H, int64_t W, int num_steps, int flags) {
  if (B <= 0 || T1 <= 0 || H <= 0 || W <= 0 || num_steps <= 0) return 0;
  const int64_t P = B * T1, field = 2 * H * W;
  // per-CTA scratch [u | m0], hand-over [u_s | m0] of the tail pairs of the dynamic schedule, ticket counter + one
  // flag per (tail pair, chunk of steps)
  if (fused_size(H, W, flags)) {
    const size_t g = (size_t)fused_grid(P, H);
    return (int64_t)(align256(sizeof(float) * g * 4 * field) + align256(16 + sizeof(int) * g * (size_t)(num_steps / kChunkSteps + 1)));
  }
  if (cluster_size(H, W, P, flags)) return (int64_t)align256((size_t)cluster_workspace_bytes(P));
  // path B: m0 (if not given) + u scratch + m/v buffer + FFT scratch
  return (int64_t)(3 * align256(sizeof(float) * (size_t)P * field) + align256((size_t)b2_fluid_workspace_bytes(P, H, W)));
}

extern "C" int b2_shoot_fwd(const b2_shoot_args* args, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!args) return B2_E_NULL;
  const b2_shoot_args& a = *args;
  if (!a.v0 || !a.u) return B2_E_NULL;
  if (a.B <= 0 || a.T1 <= 0 || a.H < 2 || a.W < 2) return B2_E_SHAPE;
  if (a.num_steps < 1 || a.num_steps > 4096 || !(a.gamma > 0.f) || a.alpha < 0.f || a.beta < 0.f || !(a.T > 0.f)) return B2_E_PARAM;
  if (a.background != B2_BG_CLAMP && a.background != B2_BG_ZERO) return B2_E_PARAM;
  if (a.sdef && !a.src) return B2_E_NULL;
  if (a.loss_terms && (!a.src || !a.tar)) return B2_E_NULL;
  if (a.S && (!a.tar || !a.moments || !a.table)) return B2_E_NULL;
  if (a.S && (a.n_sectors < 3 || a.n_sectors > kFusedMaxSectors || a.n_frames < 1)) return B2_E_PARAM;
  const int64_t P = a.B * a.T1, H = a.H, W = a.W, field = 2 * H * W;
  if (P > ((int64_t)1 << 30)) return B2_E_SHAPE;
  if (a.table_slice_stride < 0 || (a.flags & ~B2_FLAG_OPLEVEL)) return B2_E_PARAM;
  const int64_t need = b2_shoot_workspace_bytes_flags(a.B, a.T1, H, W, a.num_steps, a.flags);
  if (need <= 0) return B2_E_FFTSIZE;
  if (!workspace || workspace_bytes < need) return B2_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  // optional pair range [p0, p0 + np) of the batch (cluster path and op-level path, inference only)
  if (a.pair_count < 0 || a.pair_begin < 0 || (a.pair_count == 0 && a.pair_begin != 0)) return B2_E_PARAM;
  const bool ranged = a.pair_count > 0;
  const int64_t p0 = ranged ? a.pair_begin : 0, np = ranged ? a.pair_count : P;
  if (p0 + np > P) return B2_E_SHAPE;
  if (ranged && (a.traj || fused_size(H, W, a.flags))) return B2_E_PARAM;

  if (fused_size(H, W, a.flags)) {
    ShootParams prm;
    prm.a = a;
    prm.scratch = reinterpret_cast<float*>(workspace);
    prm.P = P;
    prm.field = field;
    const int64_t grid = fused_grid(P, H);
    // dynamic schedule (ticket counter + hand-over of the tail pairs): worth it once every CTA has at least two
    // pairs; its counter and flags sit behind the scratch and are cleared on the stream in front of the kernel
    prm.balanced = (B2_BALANCED && P >= 2 * grid && a.num_steps >= 2 * kChunkSteps) ? 1 : 0;
    if (prm.balanced)
      B2_CUDA(cudaMemsetAsync(reinterpret_cast<unsigned char*>(workspace) + align256(sizeof(float) * (size_t)grid * 4 * field),
                              0, 16 + sizeof(int) * (size_t)grid * (size_t)(a.num_steps / kChunkSteps + 1), st));
    switch ((int)H) {
      case 16: return launch_fused<16, 16, 128>(prm, grid, st);
      case 32: return launch_fused<32, 32, 256>(prm, grid, st);
      case 64: return launch_fused<64, 64, 256>(prm, grid, st);
      case 128: return launch_fused<128, 128, 1024>(prm, grid, st);
    }
    return B2_E_FFTSIZE;
  }

  if (cluster_size(H, W, np, a.flags)) {
    if (a.S && (a.n_sectors > kMaxSectors)) return B2_E_PARAM;
    return launch_shoot_cluster(a, workspace, st);
  }

  // ---- path B: op-level sequence (on the pairs [p0, p0 + np); the per-pair tensors are addressed from pair p0)
  if (a.src_slice_stride || a.tar_slice_stride) return B2_E_PARAM;   // strided volumes: fused path only
  if (a.loss_terms && (!a.sdef || !a.vel)) return B2_E_NULL;         // the op-level reduction reads both outputs
  if (b2_fluid_workspace_bytes(np, H, W) <= 0) return B2_E_FFTSIZE;
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  const size_t fbytes = align256(sizeof(float) * (size_t)P * field);
  const size_t off2 = (size_t)p0 * field, off1 = (size_t)p0 * H * W;  // first element of pair p0 in 2- / 1-plane tensors
  const float* v0p = a.v0 + off2;
  float* up = a.u + off2;
  float* velp = a.vel ? a.vel + off2 : nullptr;
  const float* m0 = v0p;
  if (!a.v0_is_momentum) {
    float* m0w = a.m0 ? a.m0 + off2 : reinterpret_cast<float*>(ws);
    if (int e = fluid_apply_impl(v0p, m0w, np, H, W, a.alpha, a.beta, a.gamma, 0, ws + 3 * fbytes,
                                 b2_fluid_workspace_bytes(np, H, W), st))
      return e;
    m0 = m0w;
  }
  float* uscr = reinterpret_cast<float*>(ws + fbytes);
  void* fws = ws + 3 * fbytes;
  const int S = a.num_steps;
  const float dt = a.T / (float)S;
  const size_t n = (size_t)P * field;                                // trajectory stride (never with a pair range)
  const float* ucur = nullptr;
  for (int s = 0; s < S; ++s) {
    // v_s goes to the trajectory (training), to `vel` at s = 0, or nowhere; m/v stay out of HBM otherwise
    float* vout = a.traj ? a.traj + (size_t)(s * 2 + 1) * n : ((s == 0 && velp) ? velp : nullptr);
    float* unext;
    if (a.traj) unext = (s + 1 < S) ? a.traj + (size_t)((s + 1) * 2) * n : up;
    else unext = (((S - (s + 1)) & 1) == 0) ? up : uscr;
    if (int e = epdiff_step_big(ucur, m0, unext, vout, fws, np, H, W, a.alpha, a.beta, a.gamma, dt, a.background, st)) return e;
    if (s == 0) {
      if (a.traj) {
        B2_CUDA(cudaMemsetAsync(a.traj, 0, sizeof(float) * n, st));
        if (a.vel) B2_CUDA(cudaMemcpyAsync(a.vel, vout, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
      }
    }
    ucur = unext;
  }
  if (a.sdef) {
    if (a.src_per_pair) {
      if (int e = b2_interp_fwd(a.src + off1, up, a.sdef + off1, np, np, np, 1, H, W, 1.f, a.background, stream)) return e;
    } else if (!ranged) {   // one source image per slice, indexed in-kernel (no repeat)
      if (int e = b2_warp_fwd(a.src, a.u, a.sdef, a.B, a.T1, 1, H, W, 1.f, a.background, stream)) return e;
    } else {                // a range may cut slices: head segment, whole slices, tail segment (one source image each)
      int64_t p = p0;
      const int64_t pend = p0 + np;
      while (p < pend) {
        const int64_t b = p / a.T1, t = p % a.T1;
        int64_t nb = 1, nt = a.T1 - t;
        if (t == 0 && pend - p >= a.T1) { nb = (pend - p) / a.T1; nt = a.T1; }
        else if (nt > pend - p) nt = pend - p;
        if (int e = b2_warp_fwd(a.src + (size_t)b * H * W, a.u + (size_t)p * field, a.sdef + (size_t)p * H * W, nb, nt, 1,
                                H, W, 1.f, a.background, stream))
          return e;
        p += nb * nt;
      }
    }
  }
  if (a.S) {
    const b2_sector_frame fr{a.table, a.table_slice_stride, a.theta0, a.clockwise};
    if (int e = strain_sector_fwd_range(a.u, a.tar, a.moments, &fr, a.S, a.counts, a.B, a.T1, H, W, a.n_sectors,
                                        a.n_frames, p0, np, st))
      return e;
  }
  if (a.loss_terms) {
    if (int e = b2_recon_loss_terms(a.sdef + off1, a.tar + off1, velp, m0, a.loss_terms + 2 * p0, np, H, W, stream)) return e;
  }
  return B2_OK;
}

extern "C" int64_t b2_sizeof_shoot_args(void) { return (int64_t)sizeof(b2_shoot_args); }

extern "C" int64_t b2_sizeof_shoot_bwd_args(void) { return (int64_t)sizeof(b2_shoot_bwd_args); }

extern "C" int64_t b2_shoot_bwd_workspace_bytes(int64_t P, int64_t H, int64_t W) {
  return b2_shoot_bwd_workspace_bytes_flags(P, H, W, 0);
}

// sized per path: resident CTAs (clusters) x 3 fields for the fused adjoints, 5 P fields + FFT scratch op-level
extern "C" int64_t b2_shoot_bwd_workspace_bytes_flags(int64_t P, int64_t H, int64_t W, int flags) {
  if (P <= 0 || H <= 0 || W <= 0) return 0;
  // per-CTA scratch (3 fields), the tail pairs' own scratch of the dynamic schedule (3 fields each), ticket + flags
  if (fused_size(H, W, flags)) {
    const size_t g = (size_t)fused_bwd_grid(P, H);
    return (int64_t)(align256(sizeof(float) * g * 6 * 2 * H * W) + align256(16 + sizeof(int) * g * kBwdMaxChunks));
  }
  if (cluster_bwd_size(H, W, P, flags)) return (int64_t)align256((size_t)cluster_bwd_workspace_bytes(P));
  const int64_t fw = b2_fluid_workspace_bytes(P, H, W);   // 0 for grids whose FFT runs in shared memory
  return (int64_t)(5 * align256(sizeof(float) * (size_t)P * 2 * H * W) + align256((size_t)fw));
}

extern "C" int b2_shoot_bwd(const float* gu, const float* gvel, const float* gm0, const float* m0, const float* traj,
                            float* gv0, int64_t P, int64_t H, int64_t W, int num_steps, float alpha, float beta,
                            float gamma, float T, int background, int v0_is_momentum, void* workspace,
                            int64_t workspace_bytes, void* stream) {
  return b2_shoot_bwd_loss(gu, gvel, gm0, nullptr, m0, traj, gv0, P, H, W, num_steps, alpha, beta, gamma, T, background,
                           v0_is_momentum, workspace, workspace_bytes, stream);
}

extern "C" int b2_shoot_bwd_loss(const float* gu, const float* gvel, const float* gm0, const float* g_reg,
                                 const float* m0, const float* traj, float* gv0, int64_t P, int64_t H, int64_t W,
                                 int num_steps, float alpha, float beta, float gamma, float T, int background,
                                 int v0_is_momentum, void* workspace, int64_t workspace_bytes, void* stream) {
  b2_shoot_bwd_args a = {};
  a.gu = gu; a.gvel = gvel; a.gm0 = gm0; a.g_reg = g_reg; a.m0 = m0; a.traj = traj; a.gv0 = gv0;
  a.P = P; a.H = H; a.W = W;
  a.num_steps = num_steps; a.background = background; a.v0_is_momentum = v0_is_momentum; a.flags = 0;
  a.alpha = alpha; a.beta = beta; a.gamma = gamma; a.T = T;
  return b2_shoot_bwd_ex(&a, workspace, workspace_bytes, stream);
}

extern "C" int b2_shoot_bwd_ex(const b2_shoot_bwd_args* args, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!args) return B2_E_NULL;
  const float *gu = args->gu, *gvel = args->gvel, *gm0 = args->gm0, *g_reg = args->g_reg, *m0 = args->m0, *traj = args->traj;
  float* gv0 = args->gv0;
  const int64_t P = args->P, H = args->H, W = args->W;
  const int num_steps = args->num_steps, background = args->background, v0_is_momentum = args->v0_is_momentum;
  const float alpha = args->alpha, beta = args->beta, gamma = args->gamma, T = args->T;
  if (!m0 || !traj || !gv0) return B2_E_NULL;
  if (P <= 0 || H < 2 || W < 2 || P > ((int64_t)1 << 30)) return B2_E_SHAPE;
  if (num_steps < 1 || !(gamma > 0.f) || !(T > 0.f) || (args->flags & ~B2_FLAG_OPLEVEL)) return B2_E_PARAM;
  if (background != B2_BG_CLAMP && background != B2_BG_ZERO) return B2_E_PARAM;
  const int64_t need = b2_shoot_bwd_workspace_bytes_flags(P, H, W, args->flags);
  if (need <= 0) return B2_E_FFTSIZE;
  if (!workspace || workspace_bytes < need) return B2_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  if (fused_size(H, W, args->flags) || cluster_bwd_size(H, W, P, args->flags)) {
    // fused adjoint: one persistent kernel; scratch = resident CTAs (clusters) x 3 fields
    ShootBwdParams prm{gu, gvel, gm0, g_reg, m0, traj, gv0, reinterpret_cast<float*>(workspace), P, 2 * H * W,
                       num_steps, v0_is_momentum, alpha, beta, gamma, T, {}};
    if (args->seed_gS || args->seed_g_sq) {
      // fused seeds: single-CTA adjoint only (the caller runs the two seed kernels itself on the other paths)
      if (H > 128) return B2_E_PARAM;
      if (!args->seed_u || !args->seed_tar || args->seed_T1 <= 0 || P % args->seed_T1 != 0) return B2_E_NULL;
      if (args->seed_gS && (!args->seed_counts || !args->seed_moments || !args->seed_table || args->seed_n_sectors < 3 ||
                            args->seed_n_sectors > kFusedMaxSectors || args->seed_n_frames < 1 || args->seed_table_slice_stride < 0))
        return B2_E_PARAM;
      if (args->seed_g_sq && !args->seed_src) return B2_E_NULL;
      prm.seed = BwdSeed{args->seed_gS, args->seed_counts, reinterpret_cast<const long long*>(args->seed_moments),
                         args->seed_table, args->seed_table_slice_stride, args->seed_theta0, args->seed_clockwise,
                         args->seed_g_sq, args->seed_u, args->seed_src, args->seed_tar, args->seed_T1,
                         args->seed_src_slice_stride, args->seed_tar_slice_stride, args->seed_n_sectors,
                         args->seed_n_frames, args->seed_src_per_pair};
    }
    switch ((int)H) {
      case 16: return launch_fused_bwd<16, 16, 128>(prm, background, st);
      case 32: return launch_fused_bwd<32, 32, 256>(prm, background, st);
      case 64: return launch_fused_bwd<64, 64, 256>(prm, background, st);
      case 128: return launch_fused_bwd<128, 128, 1024>(prm, background, st);
      case 256: return launch_shoot_cluster_bwd(prm, background, st);
    }
  }
  if (args->seed_gS || args->seed_g_sq) return B2_E_PARAM;     // fused seeds exist in the single-CTA adjoint only
  const size_t n = (size_t)P * 2 * H * W, fbytes = align256(sizeof(float) * n);
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  float* g_u = reinterpret_cast<float*>(ws);                  // dL/du_{s+1}
  float* g_uc = reinterpret_cast<float*>(ws + fbytes);        // compose part of dL/du_s
  float* g_v = reinterpret_cast<float*>(ws + 2 * fbytes);     // dL/dv_s, then dL/dm_s
  float* g_m0 = reinterpret_cast<float*>(ws + 3 * fbytes);    // accumulated dL/dm0
  float* wbuf = reinterpret_cast<float*>(ws + 4 * fbytes);    // adstar workspace: m0 o (id + u_s)
  void* fws = ws + 5 * fbytes;
  const int64_t fws_bytes = b2_fluid_workspace_bytes(P, H, W);
  const float dt = T / (float)num_steps;

  if (gu) B2_CUDA(cudaMemcpyAsync(g_u, gu, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  else B2_CUDA(cudaMemsetAsync(g_u, 0, sizeof(float) * n, st));
  if (gm0) B2_CUDA(cudaMemcpyAsync(g_m0, gm0, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  else B2_CUDA(cudaMemsetAsync(g_m0, 0, sizeof(float) * n, st));

  for (int s = num_steps - 1; s >= 0; --s) {
    const float* u_s = traj + (size_t)(s * 2) * n;
    const float* v_s = traj + (size_t)(s * 2 + 1) * n;
    // u_{s+1} = interp(u_s, v_s, -dt) - dt v_s
    if (int e = b2_compose_bwd(g_u, u_s, v_s, s > 0 ? g_uc : nullptr, g_v, P, H, W, -dt, background, stream)) return e;
    if (s == 0 && gvel) {
      if (int e = axpby(g_v, gvel, 1.f, 1.f, n, st)) return e;
    }
    // v_s = sharp(m_s), self-adjoint
    if (int e = fluid_apply_impl(g_v, g_v, P, H, W, alpha, beta, gamma, 1, fws, fws_bytes, st)) return e;
    if (s == 0) {
      // m_0 = Ad*_0 m0 = m0 exactly (u_0 = 0)
      if (int e = axpby(g_m0, g_v, 1.f, 1.f, n, st)) return e;
    } else {
      // m_s = Ad*_{u_s} m0: du -> g_u (dead after compose_bwd), dm0 accumulated into g_m0
      // (the compose part g_uc is added inside the kernel)
      if (int e = adstar_bwd_impl(g_v, u_s, m0, g_u, g_m0, wbuf, P, H, W, background, /*zero_dm0=*/false, st, g_uc)) return e;
    }
  }
  const int field = (int)(2 * H * W);
  dim3 rgrid((unsigned)((field + 255) / 256 < 64 ? (field + 255) / 256 : 64), (unsigned)(P < kMaxGridY ? P : kMaxGridY), 1);
  if (v0_is_momentum) {   // the input was m0 itself: return dL/dm0
    B2_CUDA(cudaMemcpyAsync(gv0, g_m0, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    if (g_reg) {          // + 2 g vel, vel = v_0 of the trajectory
      add_scaled_pairs_kernel<<<rgrid, 256, 0, st>>>(gv0, traj + n, g_reg, P, field);
      B2_CHECK_LAUNCH();
    }
    return B2_OK;
  }
  // dL/dv0 = flat(dL/dm0)
  if (int e = fluid_apply_impl(g_m0, gv0, P, H, W, alpha, beta, gamma, 0, fws, fws_bytes, st)) return e;
  if (g_reg) {
    add_scaled_pairs_kernel<<<rgrid, 256, 0, st>>>(gv0, m0, g_reg, P, field);
    B2_CHECK_LAUNCH();
  }
  return B2_OK;
}

extern "C" int b2_device_sm_count(int device) {
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
  return sms;
}

extern "C" int b2_shoot_cluster_occupancy(int* clusters, int* idle_sms) {
  if (!clusters || !idle_sms) return B2_E_NULL;
  int dev = 0, sms = 0;
  B2_CUDA(cudaGetDevice(&dev));
  B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int n = cluster_grid_clusters((int64_t)1 << 20);
  *clusters = n;
  *idle_sms = n > 0 ? sms - 4 * n : 0;
  return B2_OK;
}

extern "C" int b2_version(void) { return 100; }

extern "C" const char* b2_error_string(int code) {
  switch (code) {
    case B2_OK: return "ok";
    case B2_E_NULL: return "required pointer is NULL";
    case B2_E_SHAPE: return "non-positive or unsupported dimension";
    case B2_E_BCAST: return "batch sizes do not broadcast";
    case B2_E_FFTSIZE: return "H/W not a supported FFT size (square 16..128, 256x256, 64x128, 128x64, 128x256, 256x128)";
    case B2_E_PARAM: return "bad scalar parameter";
    case B2_E_WORKSPACE: return "workspace missing or too small";
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown b2lddmm error";
}

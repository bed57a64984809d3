/*
 * Plain-C restatement of the registration-to-strain forward path (TEST INFRASTRUCTURE ONLY).
 *
 * PARITY UNPINNED: like oracle/lddmm.py this restates the published lagomorph / PyCA algorithm
 *  because the reference tree holds neither the `models` package nor lagomorph
 * (/root/reference/main.py:42, README.md:15-17).  It is written independently of the torch oracle - own FFT,
 * own loops - so the two oracles cross-check each other (tests/test_oracle_cpu.py::test_c_oracle_matches_torch),
 * and it is the multi-threaded CPU baseline of bench.py (one OpenMP thread per frame-pair).
 *
 * Conventions (defaults of SURVEY.md 8c): clamp-to-edge bilinear taps (D1), one-sided differences at the image
 * edge (D2), Ad* without det (D3), expmap returns the inverse-map displacement (D4), velocity = sharp(m0) (D5),
 * strain about the frame-0 mask centroid (D6), integer-only sector classification (D7).
 *
 * Layout: fields (P, 2, H, W) fp32, component 0 along rows; cine volume (B, 1, T, H, W); P = B*(T-1) pairs,
 * Lagrangian split (/root/reference/modules/data/__init__.py:108-110).  H and W must be powers of two.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float re, im; } cpx;

/* ---- iterative radix-2 FFT, length n (power of two), stride s; sign = -1 forward, +1 inverse (unnormalised) */
static void fft1d(cpx* x, int n, int s, int sign, const cpx* tw /* tw[k] = exp(-2 pi i k / n), k < n/2 */) {
  for (int i = 1, j = 0; i < n; ++i) {               /* bit reversal */
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { cpx t = x[i * s]; x[i * s] = x[j * s]; x[j * s] = t; }
  }
  for (int len = 2; len <= n; len <<= 1) {
    const int half = len >> 1, step = n / len;
    for (int i = 0; i < n; i += len)
      for (int k = 0; k < half; ++k) {
        const cpx w = tw[k * step];
        const float wr = w.re, wi = sign < 0 ? w.im : -w.im;
        cpx* a = &x[(i + k) * s];
        cpx* b = &x[(i + k + half) * s];
        const float tr = b->re * wr - b->im * wi, ti = b->re * wi + b->im * wr;
        b->re = a->re - tr; b->im = a->im - ti;
        a->re += tr; a->im += ti;
      }
  }
}

typedef struct {
  int H, W;
  cpx *twH, *twW;          /* twiddles */
  float *c0, *s0, *c1, *s1; /* symbol LUTs over the FULL axes: 2(1-cos), sin */
  float alpha, beta, gamma;
} plan_t;

static void plan_init(plan_t* p, int H, int W, float alpha, float beta, float gamma) {
  p->H = H; p->W = W; p->alpha = alpha; p->beta = beta; p->gamma = gamma;
  p->twH = (cpx*)malloc(sizeof(cpx) * (H / 2 + 1));
  p->twW = (cpx*)malloc(sizeof(cpx) * (W / 2 + 1));
  for (int k = 0; k < H / 2; ++k) { p->twH[k].re = (float)cos(-2.0 * M_PI * k / H); p->twH[k].im = (float)sin(-2.0 * M_PI * k / H); }
  for (int k = 0; k < W / 2; ++k) { p->twW[k].re = (float)cos(-2.0 * M_PI * k / W); p->twW[k].im = (float)sin(-2.0 * M_PI * k / W); }
  p->c0 = (float*)malloc(sizeof(float) * H); p->s0 = (float*)malloc(sizeof(float) * H);
  p->c1 = (float*)malloc(sizeof(float) * W); p->s1 = (float*)malloc(sizeof(float) * W);
  for (int k = 0; k < H; ++k) { p->c0[k] = (float)(2.0 * (1.0 - cos(2.0 * M_PI * k / H))); p->s0[k] = (float)sin(2.0 * M_PI * k / H); }
  for (int k = 0; k < W; ++k) { p->c1[k] = (float)(2.0 * (1.0 - cos(2.0 * M_PI * k / W))); p->s1[k] = (float)sin(2.0 * M_PI * k / W); }
}
static void plan_free(plan_t* p) { free(p->twH); free(p->twW); free(p->c0); free(p->s0); free(p->c1); free(p->s1); }

static void fft2d(cpx* z, const plan_t* p, int sign) {
  const int H = p->H, W = p->W;
  for (int r = 0; r < H; ++r) fft1d(z + (size_t)r * W, W, 1, sign, p->twW);
  for (int c = 0; c < W; ++c) fft1d(z + c, H, W, sign, p->twH);
}

/* A.5: out = L f (inverse = 0) or L^-1 f (inverse = 1); two full complex transforms, one per real component */
static void fluid_apply(const float* f, float* out, const plan_t* p, int inverse, cpx* z0, cpx* z1) {
  const int H = p->H, W = p->W, N = H * W;
  for (int i = 0; i < N; ++i) { z0[i].re = f[i]; z0[i].im = 0.f; z1[i].re = f[N + i]; z1[i].im = 0.f; }
  fft2d(z0, p, -1); fft2d(z1, p, -1);
  const float scale = 1.0f / (float)N;
  for (int k0 = 0; k0 < H; ++k0)
    for (int k1 = 0; k1 < W; ++k1) {
      const float lam = p->gamma + p->alpha * (p->c0[k0] + p->c1[k1]);
      const float L00 = lam + p->beta * p->c0[k0], L11 = lam + p->beta * p->c1[k1];
      const float L01 = p->beta * (p->s0[k0] * p->s1[k1]);
      float a, d, b;
      if (inverse) { const float det = L00 * L11 - L01 * L01; a = L11 / det; d = L00 / det; b = -L01 / det; }
      else { a = L00; d = L11; b = L01; }
      cpx* F0 = &z0[k0 * W + k1]; cpx* F1 = &z1[k0 * W + k1];
      const cpx G0 = { scale * (a * F0->re + b * F1->re), scale * (a * F0->im + b * F1->im) };
      const cpx G1 = { scale * (b * F0->re + d * F1->re), scale * (b * F0->im + d * F1->im) };
      *F0 = G0; *F1 = G1;
    }
  fft2d(z0, p, +1); fft2d(z1, p, +1);
  for (int i = 0; i < N; ++i) { out[i] = z0[i].re; out[N + i] = z1[i].re; }
}

/* A.1: bilinear sample of plane f at (p0, p1), clamp-to-edge taps */
static inline float bilerp(const float* f, int H, int W, float p0, float p1) {
  const float f0 = floorf(p0), f1 = floorf(p1);
  const float a = p0 - f0, b = p1 - f1;
  float g0 = f0, g1 = f1;
  if (g0 < -2.f) g0 = -2.f; if (g0 > (float)(H + 1)) g0 = (float)(H + 1);
  if (g1 < -2.f) g1 = -2.f; if (g1 > (float)(W + 1)) g1 = (float)(W + 1);
  int i0 = (int)g0, j0 = (int)g1, i1 = i0 + 1, j1 = j0 + 1;
  if (i0 < 0) i0 = 0; if (i0 > H - 1) i0 = H - 1; if (i1 < 0) i1 = 0; if (i1 > H - 1) i1 = H - 1;
  if (j0 < 0) j0 = 0; if (j0 > W - 1) j0 = W - 1; if (j1 < 0) j1 = 0; if (j1 > W - 1) j1 = W - 1;
  return (((1.f - a) * (1.f - b)) * f[i0 * W + j0] + ((1.f - a) * b) * f[i0 * W + j1])
       + (a * (1.f - b)) * f[i1 * W + j0] + (a * b) * f[i1 * W + j1];
}

/* A.3: central difference, one-sided at the first/last index */
static inline float ddr(const float* f, int H, int W, int r, int c) {
  if (r == 0) return f[W + c] - f[c];
  if (r == H - 1) return f[(H - 1) * W + c] - f[(H - 2) * W + c];
  return 0.5f * (f[(r + 1) * W + c] - f[(r - 1) * W + c]);
}
static inline float ddc(const float* f, int H, int W, int r, int c) {
  (void)H;
  if (c == 0) return f[r * W + 1] - f[r * W];
  if (c == W - 1) return f[r * W + W - 1] - f[r * W + W - 2];
  return 0.5f * (f[r * W + c + 1] - f[r * W + c - 1]);
}

/* D7: integer-only sector of direction (dr, dc) in the slice's sector frame: table[2k] = Q20 sin, table[2k+1] =
 * Q20 cos of theta0 + 2 pi k / n (start angle of the reference's mesh, DENSE_utils.py:198); clockwise == 0 numbers
 * the sectors the other way round (DENSE_utils.py:201-204): k -> n-1-k */
static int classify(long long dr, long long dc, const int32_t* tab, int n, double theta0, int clockwise) {
  if (dr == 0 && dc == 0) return -1;
  double th = atan2((double)dr, (double)dc) - theta0;
  th -= 2.0 * M_PI * floor(th / (2.0 * M_PI));
  int k = (int)floor(th / (2.0 * M_PI / n));
  if (k < 0) k = 0; if (k > n - 1) k = n - 1;
  for (int it = 0; it < n; ++it) {
    const int k1 = (k + 1 == n) ? 0 : k + 1;
    const long long lo = (long long)tab[2 * k + 1] * dr - (long long)tab[2 * k] * dc;
    const long long hi = (long long)tab[2 * k1 + 1] * dr - (long long)tab[2 * k1] * dc;
    if (lo < 0) k = (k == 0) ? n - 1 : k - 1;
    else if (hi >= 0) k = k1;
    else break;
  }
  return clockwise ? k : n - 1 - k;
}

void b2o_sector_table_rotated(int n, double theta0, int32_t* tab) {
  for (int k = 0; k < n; ++k) {
    const double ang = theta0 + 2.0 * M_PI * (double)k / (double)n;
    tab[2 * k] = (int32_t)llrint(1048576.0 * sin(ang));
    tab[2 * k + 1] = (int32_t)llrint(1048576.0 * cos(ang));
  }
}

void b2o_sector_table(int n, int32_t* tab) { b2o_sector_table_rotated(n, 0.0, tab); }

/* Whole forward path for a batch of slices; theta0 (B doubles) / clockwise (B ints): per-slice sector frame, NULL = 0 / 1.  Outputs: m0, vel, u (P,2,H,W); sdef (P,1,H,W);
 * S (B,1,n_sectors,n_frames).  Returns 0, or -1 on bad sizes / allocation failure. */
int b2o_forward_volume(const float* v0, const float* vol, int B, int T, int H, int W, int num_steps,
                       float alpha, float beta, float gamma, int n_sectors, int n_frames,
                       float* m0_out, float* vel_out, float* u_out, float* sdef_out, float* S_out, int nthreads,
                       const double* theta0, const int* clockwise) {
  if (B < 1 || T < 2 || H < 2 || W < 2 || (H & (H - 1)) || (W & (W - 1)) || num_steps < 1 || n_sectors < 3) return -1;
  const int T1 = T - 1, N = H * W, P = B * T1;
  const float dt = 1.0f / (float)num_steps;
  int32_t* tab = (int32_t*)malloc(sizeof(int32_t) * 2 * n_sectors * (size_t)B);   /* one rotated table per slice */
  if (!tab) return -1;
  for (int b = 0; b < B; ++b) b2o_sector_table_rotated(n_sectors, theta0 ? theta0[b] : 0.0, tab + (size_t)b * 2 * n_sectors);
  memset(S_out, 0, sizeof(float) * (size_t)B * n_sectors * n_frames);
  int failed = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    plan_t pl;
    plan_init(&pl, H, W, alpha, beta, gamma);
    cpx* z0 = (cpx*)malloc(sizeof(cpx) * N);
    cpx* z1 = (cpx*)malloc(sizeof(cpx) * N);
    float* m = (float*)malloc(sizeof(float) * 2 * N);
    float* v = (float*)malloc(sizeof(float) * 2 * N);
    float* un = (float*)malloc(sizeof(float) * 2 * N);
    double* sums = (double*)malloc(sizeof(double) * n_sectors);
    int* cnts = (int*)malloc(sizeof(int) * n_sectors);
    if (!z0 || !z1 || !m || !v || !un || !sums || !cnts) {
#pragma omp atomic write
      failed = 1;
    }
#pragma omp for schedule(dynamic, 1)
    for (int p = 0; p < P; ++p) {
      if (failed) continue;
      const int b = p / T1, t = p % T1;
      const float* src = vol + (size_t)b * T * N;             /* frame 0 */
      const float* tar = vol + ((size_t)b * T + t + 1) * N;   /* frame t+1 */
      float* m0 = m0_out + (size_t)p * 2 * N;
      float* u = u_out + (size_t)p * 2 * N;
      fluid_apply(v0 + (size_t)p * 2 * N, m0, &pl, 0, z0, z1);            /* m0 = flat(v0) */
      memset(u, 0, sizeof(float) * 2 * N);
      for (int s = 0; s < num_steps; ++s) {
        /* A.4: m = (I + Du)^T (m0 o (id + u)) */
        for (int r = 0; r < H; ++r)
          for (int c = 0; c < W; ++c) {
            const int i = r * W + c;
            const float p0 = (float)r + u[i], p1 = (float)c + u[N + i];
            const float w0 = bilerp(m0, H, W, p0, p1), w1 = bilerp(m0 + N, H, W, p0, p1);
            const float d00 = ddr(u, H, W, r, c), d10 = ddr(u + N, H, W, r, c);
            const float d01 = ddc(u, H, W, r, c), d11 = ddc(u + N, H, W, r, c);
            m[i] = w0 + (d00 * w0 + d10 * w1);
            m[N + i] = w1 + (d01 * w0 + d11 * w1);
          }
        fluid_apply(m, v, &pl, 1, z0, z1);                                /* v = sharp(m) */
        if (s == 0) memcpy(vel_out + (size_t)p * 2 * N, v, sizeof(float) * 2 * N);
        /* A.6: u <- interp(u, v, -dt) - dt v */
        for (int r = 0; r < H; ++r)
          for (int c = 0; c < W; ++c) {
            const int i = r * W + c;
            const float p0 = (float)r - dt * v[i], p1 = (float)c - dt * v[N + i];
            un[i] = bilerp(u, H, W, p0, p1) - dt * v[i];
            un[N + i] = bilerp(u + N, H, W, p0, p1) - dt * v[N + i];
          }
        memcpy(u, un, sizeof(float) * 2 * N);
      }
      /* deformed source */
      float* sd = sdef_out + (size_t)p * N;
      for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
          const int i = r * W + c;
          sd[i] = bilerp(src, H, W, (float)r + u[i], (float)c + u[N + i]);
        }
      /* A.7 / A.8: strain matrix column t of slice b */
      long long cnt = 0, sx = 0, sy = 0;
      for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c)
          if (src[r * W + c] > 0.5f) { cnt += 1; sx += r; sy += c; }
      const float ctr0 = cnt > 0 ? (float)((double)sx / (double)cnt) : (float)((H - 1) * 0.5);
      const float ctr1 = cnt > 0 ? (float)((double)sy / (double)cnt) : (float)((W - 1) * 0.5);
      for (int k = 0; k < n_sectors; ++k) { sums[k] = 0.0; cnts[k] = 0; }
      for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
          const int i = r * W + c;
          if (!(tar[i] > 0.5f)) continue;
          const int k = classify(cnt * r - sx, cnt * c - sy, tab + (size_t)b * 2 * n_sectors, n_sectors,
                                 theta0 ? theta0[b] : 0.0, clockwise ? clockwise[b] : 1);
          if (k < 0) continue;
          const float G00 = 1.f + ddr(u, H, W, r, c), G01 = ddc(u, H, W, r, c);
          const float G10 = ddr(u + N, H, W, r, c), G11 = 1.f + ddc(u + N, H, W, r, c);
          const float det = G00 * G11 - G01 * G10;
          const float n0 = ((float)r + u[i]) - ctr0, n1 = ((float)c + u[N + i]) - ctr1;
          const float rad2 = n0 * n0 + n1 * n1;
          if (rad2 < 1e-12f || fabsf(det) < 1e-6f) continue;
          const float e0 = -n1, e1 = n0;
          const float t0 = G11 * e0 - G01 * e1, t1 = G00 * e1 - G10 * e0;
          const float ecc = 0.5f * ((t0 * t0 + t1 * t1) / (rad2 * det * det) - 1.f);
          sums[k] += (double)ecc;
          cnts[k] += 1;
        }
      for (int k = 0; k < n_sectors; ++k) {
        const float val = (float)(sums[k] / (double)(cnts[k] > 0 ? cnts[k] : 1));
        float* row = S_out + ((size_t)b * n_sectors + k) * n_frames;
        if (t < n_frames) row[t] = val;
        if (t == T1 - 1) for (int tt = T1; tt < n_frames; ++tt) row[tt] = val;
      }
    }
    free(z0); free(z1); free(m); free(v); free(un); free(sums); free(cnts);
    plan_free(&pl);
  }
  free(tab);
  return failed ? -1 : 0;
}

int b2o_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

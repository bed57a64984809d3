// Stand-alone use of the C ABI (include/b2lddmm.h) without PyTorch: cudaMalloc'd buffers, one call to the fused
// shooting kernel, results read back and summarised.  Build and run (B200):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I include examples/c_abi_demo.cu \
//        -L <pkg> -lb2lddmm -Xlinker -rpath=<pkg> -o build/c_abi_demo && build/c_abi_demo
// Prints a JSON line with checksums that tests/test_gpu_parity.py::test_c_abi_standalone compares with the
// same computation made through the Python surface.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "b2lddmm.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 2; } } while (0)
#define B2(x) do { int rc = (x); if (rc != 0) { fprintf(stderr, "%s: %s\n", #x, b2_error_string(rc)); return 3; } } while (0)

int main(int argc, char** argv) {
  const int64_t B = 2, T1 = 3, H = 64, W = 64, N = H * W, P = B * T1;
  const int S = 4, n_sectors = 126, n_frames = 40;
  // deterministic synthetic inputs: discs as masks, a smooth analytic velocity field
  std::vector<float> vol((size_t)B * (T1 + 1) * N), v0((size_t)P * 2 * N);
  for (int64_t b = 0; b < B; ++b)
    for (int64_t t = 0; t <= T1; ++t)
      for (int64_t r = 0; r < H; ++r)
        for (int64_t c = 0; c < W; ++c) {
          const double dr = r - 31.5 - b, dc = c - 31.5 + b, rad = sqrt(dr * dr + dc * dc);
          vol[((b * (T1 + 1) + t) * H + r) * W + c] = (rad >= 10.0 - t && rad <= 20.0 - 0.5 * t) ? 1.f : 0.f;
        }
  for (int64_t p = 0; p < P; ++p)
    for (int64_t r = 0; r < H; ++r)
      for (int64_t c = 0; c < W; ++c) {
        const double a = 2.0 * M_PI * r / H, bb = 2.0 * M_PI * c / W;
        v0[((p * 2 + 0) * H + r) * W + c] = (float)(1.5 * sin(a + 0.3 * p) * cos(bb));
        v0[((p * 2 + 1) * H + r) * W + c] = (float)(1.5 * cos(a) * sin(bb - 0.2 * p));
      }
  float *d_vol, *d_v0, *d_m0, *d_vel, *d_u, *d_sdef, *d_S;
  int64_t* d_mom; int32_t *d_tab, *d_cnt; void* d_ws;
  CK(cudaMalloc(&d_vol, vol.size() * 4)); CK(cudaMalloc(&d_v0, v0.size() * 4));
  CK(cudaMalloc(&d_m0, v0.size() * 4)); CK(cudaMalloc(&d_vel, v0.size() * 4)); CK(cudaMalloc(&d_u, v0.size() * 4));
  CK(cudaMalloc(&d_sdef, (size_t)P * N * 4)); CK(cudaMalloc(&d_S, (size_t)B * n_sectors * n_frames * 4));
  CK(cudaMalloc(&d_mom, B * 3 * 8)); CK(cudaMalloc(&d_tab, 2 * n_sectors * 4)); CK(cudaMalloc(&d_cnt, B * n_sectors * T1 * 4));
  CK(cudaMemcpy(d_vol, vol.data(), vol.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_v0, v0.data(), v0.size() * 4, cudaMemcpyHostToDevice));
  std::vector<int32_t> tab(2 * n_sectors);
  B2(b2_sector_table_host(n_sectors, tab.data()));
  CK(cudaMemcpy(d_tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  // frame-0 masks of the slices, gathered densely for the moments kernel
  float* d_mask0; CK(cudaMalloc(&d_mask0, (size_t)B * N * 4));
  for (int64_t b = 0; b < B; ++b)
    CK(cudaMemcpyAsync(d_mask0 + b * N, d_vol + b * (T1 + 1) * N, N * 4, cudaMemcpyDeviceToDevice, st));
  B2(b2_mask_moments(d_mask0, d_mom, B, H, W, st));
  b2_shoot_args a = {};
  a.v0 = d_v0; a.src = d_vol; a.tar = d_vol + N; a.moments = d_mom; a.table = d_tab;
  a.m0 = d_m0; a.vel = d_vel; a.u = d_u; a.sdef = d_sdef; a.S = d_S; a.counts = d_cnt; a.traj = nullptr;
  a.B = B; a.T1 = T1; a.H = H; a.W = W;
  a.src_slice_stride = (T1 + 1) * N; a.tar_slice_stride = (T1 + 1) * N;      // read the cine volume in place
  a.num_steps = S; a.src_per_pair = 0; a.v0_is_momentum = 0;
  a.n_sectors = n_sectors; a.n_frames = n_frames; a.background = B2_BG_CLAMP;
  a.alpha = 1.0f; a.beta = 0.1f; a.gamma = 0.05f; a.T = 1.0f;
  const int64_t wsb = b2_shoot_workspace_bytes(B, T1, H, W, S);
  CK(cudaMalloc(&d_ws, wsb));
  B2(b2_shoot_fwd(&a, d_ws, wsb, st));
  CK(cudaStreamSynchronize(st));
  std::vector<float> u(v0.size()), sdef((size_t)P * N), Sm((size_t)B * n_sectors * n_frames);
  CK(cudaMemcpy(u.data(), d_u, u.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(sdef.data(), d_sdef, sdef.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(Sm.data(), d_S, Sm.size() * 4, cudaMemcpyDeviceToHost));
  double su = 0, ss = 0, sS = 0;
  for (float x : u) su += fabs(x);
  for (float x : sdef) ss += x;
  for (float x : Sm) sS += fabs(x);
  printf("{\"b2_version\": %d, \"sum_abs_u\": %.9e, \"sum_sdef\": %.9e, \"sum_abs_S\": %.9e}\n", b2_version(), su, ss, sS);
  if (argc > 1) {   // dump raw outputs for an exact comparison
    FILE* f = fopen(argv[1], "wb");
    if (!f) return 4;
    fwrite(v0.data(), 4, v0.size(), f); fwrite(vol.data(), 4, vol.size(), f);
    fwrite(u.data(), 4, u.size(), f); fwrite(sdef.data(), 4, sdef.size(), f); fwrite(Sm.data(), 4, Sm.size(), f);
    fclose(f);
  }
  return 0;
}

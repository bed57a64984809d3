// Parameter blocks shared by the single-CTA (shoot.cu) and cluster (shoot_cluster.cu) shooting kernels.
#pragma once
#include "common.cuh"

namespace b2 {

// Optional seeds of dL/du^S computed in the prologue of the fused adjoint kernel instead of by separate kernels
// (b2_strain_sector_bwd_ex + b2_warp_sqerr_bwd) through a (P,2,H,W) gradient image: the adjoint of the strain-matrix
// reduction (gS != nullptr) and of the squared-error term of the reconstruction loss (g_sq != nullptr).
struct BwdSeed {
  const float* gS;            // (B,1,n_sectors,n_frames) dL/dS
  const int32_t* counts;      // (B,n_sectors,T1) member counts of the forward
  const long long* moments;   // (B,3)
  const int32_t* table;       // sector frame (b2_sector_frame)
  long long table_slice_stride;
  const float* theta0;
  const int32_t* clockwise;
  const float* g_sq;          // (P) dL/d(sum (tar - Sdef)^2)
  const float* uS;            // (P,2,H,W) final displacement u^S
  const float* src;
  const float* tar;
  long long T1, src_slice_stride, tar_slice_stride;
  int n_sectors, n_frames, src_per_pair;
};

// Fused EPDiff adjoint (reverse sweep over the saved trajectory, b2_shoot_bwd_ex)
struct ShootBwdParams {
  const float* gu;      // dL/du^S   (P,2,H,W) or nullptr
  const float* gvel;    // dL/dvel   or nullptr
  const float* gm0;     // explicit dL/dm0 or nullptr
  const float* g_reg;   // (P) dL/d(sum vel . m0) or nullptr: closed-form 2 g m0 (2 g vel) added to the result
  const float* m0;
  const float* traj;    // (S, 2, P, 2, H, W)
  float* gv0;
  float* scratch;       // per CTA: [G ping | G pong | dL/dm0] (w = m0 o (id + u_s) reuses the dead G buffer);
                        // per cluster (256x256): [G ping | G pong | dL/dm0 | row-neighbour products]
  int64_t P, field;
  int num_steps, v0_is_momentum;
  float alpha, beta, gamma, T;
  BwdSeed seed;         // all-null = none
  // dynamic ticket schedule of the single-CTA kernel (shoot.cu): null ticket = static round-robin over the pairs
  unsigned long long* ticket;   // zeroed per launch
  int* flags;                   // flags[i * n_chunks + c] = chunk c of tail pair i done (zeroed per launch)
  int chunk_steps;              // adjoint steps per chunk of a tail pair
};

}  // namespace b2

// Parameter blocks shared by the single-CTA (shoot.cu) and cluster (shoot_cluster.cu) shooting kernels.
#pragma once
#include "common.cuh"

namespace b2 {

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
};

}  // namespace b2

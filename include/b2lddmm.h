/*
 * b2lddmm.h - C ABI of the B200-native registration-to-strain hot path.
 *
 * This is the drop-in boundary.  The reference reaches
 * this path through the Python operator surface of the third-party `lagomorph`
 * package (imported at /root/reference/modules/trainer/joint_registration_strainmat_LMA.py:5,
 * reg_trainer.py:4) whose native layer is the pybind11 extension `lagomorph_ext`;
 * neither is vendored in the reference tree.  Each entry point below names the
 * lagomorph / lagomorph_ext function it replaces and the reference call site
 * that consumes its result.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *    name ends in `_host`; tensors are contiguous fp32, layout (P, C, H, W),
 *    vector fields C = 2 with component 0 along rows (H), 1 along columns (W),
 *    displacements in pixels.
 *  - batch broadcast: PI / Pu are either 1 or P (lagomorph broadcasts I or u).
 *  - `stream` is a cudaStream_t passed as void*; all work is stream ordered;
 *    no call allocates, frees or synchronises; scratch comes from the caller
 *    (`*_workspace_bytes` queries).
 *  - return value: 0 = ok, < 0 = invalid argument (B2_E_*), > 0 = cudaError_t.
 *    Nothing throws.  b2_error_string() decodes either range.
 *  - background: 0 = clamp-to-edge (D1 default), 1 = zero.
 *  - FFT sizes: H and W powers of two in [16, 256] (fluid metric, shooting).
 */
#ifndef B2LDDMM_H
#define B2LDDMM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_OK 0
#define B2_E_NULL (-1)       /* required pointer is NULL */
#define B2_E_SHAPE (-2)      /* non-positive or unsupported dimension */
#define B2_E_BCAST (-3)      /* batch sizes do not broadcast */
#define B2_E_FFTSIZE (-4)    /* H or W not a supported power of two */
#define B2_E_PARAM (-5)      /* bad scalar parameter (gamma <= 0, n_sectors, flags...) */
#define B2_E_WORKSPACE (-6)  /* workspace missing or too small */

#define B2_BG_CLAMP 0
#define B2_BG_ZERO 1

/* Path selection is explicit (no environment variables are read anywhere in the library): `flags` of
 * b2_shoot_args / b2_shoot_bwd_args and of the *_flags workspace queries. */
#define B2_FLAG_OPLEVEL 1    /* run the op-level kernel sequence (path B) even where a fused kernel exists */

/* ---- sector frame of a slice (DENSE_utils.py:196-204 of the reference: spl2patchSA) ----
 * The reference's 126-sector mesh starts at theta0 = arctan2(PositionB - PositionA) and runs clockwise or
 * counter-clockwise per subject.  With theta = atan2(d_row, d_col) about the frame-0 mask centroid:
 *   clockwise != 0 (default):  sector = floor(((theta - theta0) mod 2 pi) / (2 pi / n))
 *   clockwise == 0          :  sector = n - 1 - (that index)
 * Integer-exact: `table` holds the Q20 boundary directions ALREADY ROTATED by theta0 on the host
 * (b2_sector_table_rotated_host), one (n_sectors,2) table per slice `table_slice_stride` int32 apart
 * (0 = one table shared by all slices); `theta0` (B floats, device) only seeds the search and may be NULL
 * when every table is unrotated; `clockwise` (B int32, device) NULL = all clockwise. */
typedef struct b2_sector_frame {
  const int32_t* table;
  int64_t table_slice_stride;
  const float* theta0;
  const int32_t* clockwise;
} b2_sector_frame;

int b2_version(void);
const char* b2_error_string(int code);

/* ---- lagomorph.interp  (lagomorph_ext.interp_forward / interp_backward) ----
 * out(p,c,x) = I(p,c, x + dt*u(p,:,x)), bilinear.  Consumer: 'deformed_source',
 * joint_registration_strainmat_LMA.py:315; reg_trainer.py:225. */
int b2_interp_fwd(const float* I, const float* u, float* out,
                  int64_t P, int64_t PI, int64_t Pu, int64_t C, int64_t H, int64_t W,
                  float dt, int background, void* stream);
/* dI (PI,C,H,W) and du (Pu,2,H,W) may each be NULL (gradient not needed).
 * dI is zero-filled by the call and accumulated with float atomics. */
int b2_interp_bwd(const float* gout, const float* I, const float* u, float* dI, float* du,
                  int64_t P, int64_t PI, int64_t Pu, int64_t C, int64_t H, int64_t W,
                  float dt, int background, void* stream);

/* ---- warp of frames and masks: Sdef = lagomorph.interp(src, u) with src (B,C,H,W) shared by
 * the T1 frame-pairs of its slice; u (B*T1,2,H,W); out (B*T1,C,H,W).  Replaces the
 * src.repeat(T1) + interp of the reference path (modules/data/__init__.py:109,
 * joint_registration_strainmat_LMA.py:315).  dsrc (B,C,H,W) / du may be NULL. */
int b2_warp_fwd(const float* src, const float* u, float* out, int64_t B, int64_t T1, int64_t C,
                int64_t H, int64_t W, float dt, int background, void* stream);
int b2_warp_bwd(const float* gout, const float* src, const float* u, float* dsrc, float* du,
                int64_t B, int64_t T1, int64_t C, int64_t H, int64_t W, float dt, int background,
                void* stream);

/* ---- lagomorph.splat ----  transpose of interp in I; wout (P,1,H,W) optional. */
int b2_splat_fwd(const float* J, const float* u, float* out, float* wout,
                 int64_t P, int64_t PJ, int64_t Pu, int64_t C, int64_t H, int64_t W,
                 float dt, int background, void* stream);

/* ---- lagomorph.compose_disp_vel ----  out = interp(u, v, dt) + dt*v,  all (P,2,H,W). */
int b2_compose_fwd(const float* u, const float* v, float* out,
                   int64_t P, int64_t H, int64_t W, float dt, int background, void* stream);
int b2_compose_bwd(const float* gout, const float* u, const float* v, float* du, float* dv,
                   int64_t P, int64_t H, int64_t W, float dt, int background, void* stream);

/* ---- lagomorph.jacobian_times_vectorfield ----
 * out = (displacement*I + Dv) w, or its transpose applied to w; central
 * differences, one-sided at the image edge (D2). */
int b2_jtv_fwd(const float* v, const float* w, float* out, int64_t P, int64_t H, int64_t W,
               int displacement, int transpose, void* stream);
int b2_jtv_bwd(const float* gout, const float* v, const float* w, float* dv, float* dw,
               int64_t P, int64_t H, int64_t W, int displacement, int transpose, void* stream);

/* ---- lagomorph.Ad_star ----  m = (I + Du)^T (m0 o (id + u)).
 * bwd needs a (P,2,H,W) float workspace (holds m0 o (id+u)). dm0 is zero-filled. */
int b2_adstar_fwd(const float* u, const float* m0, float* out,
                  int64_t P, int64_t H, int64_t W, int background, void* stream);
int b2_adstar_bwd(const float* gout, const float* u, const float* m0, float* du, float* dm0,
                  float* workspace, int64_t P, int64_t H, int64_t W, int background, void* stream);

/* ---- lagomorph.FluidMetric.flat / .sharp  (lagomorph_ext.fluid_operator + FFT) ----
 * L = gamma*I - alpha*Lap - beta*grad div on the periodic grid, applied in the
 * Fourier domain by an in-kernel shared-memory FFT; inverse != 0 solves (sharp).
 * Self-adjoint: the backward pass is the same call on the incoming gradient.
 * May run in place (f == out). */
int64_t b2_fluid_workspace_bytes(int64_t P, int64_t H, int64_t W);
int b2_fluid_apply(const float* f, float* out, int64_t P, int64_t H, int64_t W,
                   float alpha, float beta, float gamma, int inverse,
                   void* workspace, int64_t workspace_bytes, void* stream);

/* ---- sectors ([SPEC]; count/order pinned by DENSE_utils.py:177-295, affine.py:52-87) ----
 * Host helper: Q20 boundary directions (row, col) for n_sectors wedges. */
int b2_sector_table_host(int n_sectors, int32_t* table_host /* 2*n_sectors */);
/* Same with every boundary rotated by theta0 (radians): b_k = round(2^20 (sin, cos)(theta0 + 2 pi k / n)). */
int b2_sector_table_rotated_host(int n_sectors, double theta0, int32_t* table_host /* 2*n_sectors */);
/* moments (B,3) int64 = {count, sum(row), sum(col)} of mask0 > 0.5; zero-filled by the call. */
int b2_mask_moments(const float* mask0, int64_t* moments, int64_t B, int64_t H, int64_t W,
                    void* stream);
int b2_sector_map_i32(const int64_t* moments, const int32_t* table, int32_t* sector,
                      int64_t B, int64_t H, int64_t W, int n_sectors, void* stream);
/* per-slice sector frame (theta0 / direction); frame->table must not be NULL */
int b2_sector_map_i32_ex(const int64_t* moments, const b2_sector_frame* frame, int32_t* sector,
                         int64_t B, int64_t H, int64_t W, int n_sectors, void* stream);

/* ---- strain matrix ([SPEC] SURVEY.md A.7/A.8) ----
 * u (B,T1,2,H,W), tar (B,T1,H,W) masks, moments (B,3).  S (B,1,n_sectors,n_frames):
 * per-sector mean Ecc of frame t in column t, columns >= T1 edge-padded, columns
 * beyond n_frames cropped (align_n_frames_to, DENSE_IO_utils.py:26-46).
 * counts (B,n_sectors,T1) int32 is optional output (needed by the backward). */
int b2_strain_sector_fwd(const float* u, const float* tar, const int64_t* moments,
                         const int32_t* table, float* S, int32_t* counts,
                         int64_t B, int64_t T1, int64_t H, int64_t W,
                         int n_sectors, int n_frames, void* stream);
int b2_strain_sector_bwd(const float* gS, const float* u, const float* tar,
                         const int64_t* moments, const int32_t* table, const int32_t* counts,
                         float* du, int64_t B, int64_t T1, int64_t H, int64_t W,
                         int n_sectors, int n_frames, void* stream);
/* same two with a per-slice sector frame */
int b2_strain_sector_fwd_ex(const float* u, const float* tar, const int64_t* moments,
                            const b2_sector_frame* frame, float* S, int32_t* counts,
                            int64_t B, int64_t T1, int64_t H, int64_t W,
                            int n_sectors, int n_frames, void* stream);
int b2_strain_sector_bwd_ex(const float* gS, const float* u, const float* tar,
                            const int64_t* moments, const b2_sector_frame* frame, const int32_t* counts,
                            float* du, int64_t B, int64_t T1, int64_t H, int64_t W,
                            int n_sectors, int n_frames, void* stream);

/* ---- fused geodesic shooting: flat + S x EPDiff_step (+ warp + strain) ----
 * Replaces m0 = metric.flat(v0); u = lagomorph.expmap(metric, m0, T, num_steps);
 * Sdef = lagomorph.interp(src, u)  and the strain reduction, i.e. the body of
 * forward_volume (call site joint_registration_strainmat_LMA.py:307).
 *
 * v0 (P,2,H,W) with P = B*T1 ordered slice-major; src (B,1,H,W) frame-0 images
 * (broadcast over the T1 pairs of a slice; src_per_pair != 0 means src is (P,1,H,W));
 * tar (P,1,H,W).  Outputs (each may be NULL except u): m0 = flat(v0) 'momentum',
 * vel = sharp(m0) 'velocity', u = u^S 'displacement', sdef = interp(src,u)
 * 'deformed_source', S strain matrix (B,1,n_sectors,n_frames; needs moments+table+tar),
 * traj (num_steps, 2, P, 2, H, W): step-major (u_s, v_s) of every step, for the adjoint. */
typedef struct b2_shoot_args {
  const float* v0;
  const float* src;
  const float* tar;
  const int64_t* moments;
  const int32_t* table;
  float* m0;
  float* vel;
  float* u;
  float* sdef;
  float* S;
  int32_t* counts;
  float* traj;
  int64_t B, T1, H, W;
  int64_t src_slice_stride;  /* elements between the (first) source images of consecutive slices; 0 = dense.
                                With src_per_pair the source of pair (b,t) is src + b*stride + t*H*W. */
  int64_t tar_slice_stride;  /* elements between the first target frames of consecutive slices; 0 = T1*H*W.
                                With both set to T*H*W, src = vol and tar = vol + H*W read a (B,1,T,H,W) cine
                                volume in place: no pair construction, no copies (fused path only). */
  int32_t num_steps;
  int32_t src_per_pair;
  int32_t v0_is_momentum;  /* != 0: `v0` already holds m0 (lagomorph.expmap(metric, m0)); flat is skipped,
                              the m0 output is not written and vel = sharp(m0) */
  int32_t n_sectors, n_frames;
  int32_t background;
  float alpha, beta, gamma, T;
  float* loss_terms;       /* optional output (P,2): per frame-pair {sum_x (tar - sdef)^2, sum_x vel . m0}, the two
                              reductions of RegistrationReconstructionLoss (registration_losses.py:25-26) taken
                              inside the shooting kernel (fixed summation order: bitwise reproducible).  Needs
                              src and tar; on the op-level path (rectangular grids) also the sdef and vel outputs. */
  /* sector frame of every slice (see b2_sector_frame): `table` above is read with this stride; all optional */
  int64_t table_slice_stride;
  const float* theta0;
  const int32_t* clockwise;
  int32_t flags;           /* B2_FLAG_* */
  int32_t reserved_;
  /* Optional pair range: with pair_count > 0 only the frame-pairs [pair_begin, pair_begin + pair_count) of the B*T1
   * pairs are processed; every pointer still addresses the whole batch and the outputs of the other pairs are left
   * untouched (a slice may be cut anywhere: its strain-matrix columns are written per pair).  Lets two launches
   * share one batch at pair granularity - the 256x256 cluster kernel and the op-level sequence on the SMs the
   * clusters cannot occupy.  Supported on the 256x256 cluster path and the op-level path, without traj;
   * B2_E_PARAM elsewhere.  0 / 0 = all pairs. */
  int64_t pair_begin, pair_count;
} b2_shoot_args;
int64_t b2_sizeof_shoot_args(void);   /* ABI check for bindings that mirror the struct */

int64_t b2_shoot_workspace_bytes(int64_t B, int64_t T1, int64_t H, int64_t W, int num_steps);   /* flags = 0 */
int64_t b2_shoot_workspace_bytes_flags(int64_t B, int64_t T1, int64_t H, int64_t W, int num_steps, int flags);
int b2_shoot_fwd(const b2_shoot_args* args, void* workspace, int64_t workspace_bytes, void* stream);

/* Adjoint of b2_shoot_fwd through expmap and flat: given gu = dL/du^S (NULL = 0),
 * gvel = dL/dvel, gm0 = dL/dm0 (explicit dependence, e.g. sum(v*m); NULL = 0),
 * returns gv0 = dL/dv0 (P,2,H,W).  Needs traj from the forward. */
int64_t b2_shoot_bwd_workspace_bytes(int64_t P, int64_t H, int64_t W);
int b2_shoot_bwd(const float* gu, const float* gvel, const float* gm0, const float* m0,
                 const float* traj, float* gv0, int64_t P, int64_t H, int64_t W,
                 int num_steps, float alpha, float beta, float gamma, float T, int background,
                 int v0_is_momentum /* != 0: return dL/dm0 instead of dL/dv0 */,
                 void* workspace, int64_t workspace_bytes, void* stream);
/* Same, plus the gradient of the regularisation term taken in closed form: with g_reg (P) = dL/d(sum_x vel . m0)
 * per pair (NULL = none), d<sharp(m0), m0>/dm0 = 2 vel, hence gv0 += 2 g_reg[p] m0 (or, for v0_is_momentum,
 * gm0 += 2 g_reg[p] vel with vel = v_0 of the trajectory) - no elementwise seed tensors, no extra FFT. */
int b2_shoot_bwd_loss(const float* gu, const float* gvel, const float* gm0, const float* g_reg,
                      const float* m0, const float* traj, float* gv0, int64_t P, int64_t H, int64_t W,
                      int num_steps, float alpha, float beta, float gamma, float T, int background,
                      int v0_is_momentum, void* workspace, int64_t workspace_bytes, void* stream);

/* Struct form of the same call with explicit path flags.  Fused adjoint kernels exist for square grids of
 * 16..128 (one CTA per frame-pair) and 256x256 (one 4-CTA cluster per frame-pair); other sizes, B2_FLAG_OPLEVEL,
 * or a device that cannot co-schedule the cluster run the op-level sweep.  The workspace is sized per path:
 * resident CTAs x 3 fields (fused), 5 P fields + FFT scratch (op-level). */
typedef struct b2_shoot_bwd_args {
  const float* gu;      /* dL/du^S (P,2,H,W) or NULL */
  const float* gvel;    /* dL/dvel or NULL */
  const float* gm0;     /* explicit dL/dm0 or NULL */
  const float* g_reg;   /* (P) dL/d(sum vel . m0) or NULL */
  const float* m0;
  const float* traj;
  float* gv0;
  int64_t P, H, W;
  int32_t num_steps, background, v0_is_momentum, flags;
  float alpha, beta, gamma, T;
  /* Optional fused seeds of dL/du^S (single-CTA adjoint, square grids up to 128x128; B2_E_PARAM elsewhere - run
   * b2_strain_sector_bwd_ex / b2_warp_sqerr_bwd and pass `gu` instead): the adjoint of the strain-matrix reduction
   * (seed_gS = dL/dS (B,1,n_sectors,n_frames), with the forward's counts, moments and sector frame) and of the squared
   * error sum_x (tar - interp(src, u^S))^2 (seed_g_sq (P) = its upstream gradient per pair) are taken in the prologue
   * of the adjoint kernel from u^S, src and tar (addressing as in b2_shoot_args), added to `gu` when that is given.
   * No (P,2,H,W) gradient image, no extra launches. */
  const float* seed_gS;
  const int32_t* seed_counts;
  const int64_t* seed_moments;
  const int32_t* seed_table;
  int64_t seed_table_slice_stride;
  const float* seed_theta0;
  const int32_t* seed_clockwise;
  const float* seed_g_sq;
  const float* seed_u;
  const float* seed_src;
  const float* seed_tar;
  int64_t seed_T1, seed_src_slice_stride, seed_tar_slice_stride;
  int32_t seed_n_sectors, seed_n_frames, seed_src_per_pair, seed_reserved_;
} b2_shoot_bwd_args;
int64_t b2_sizeof_shoot_bwd_args(void);
int64_t b2_shoot_bwd_workspace_bytes_flags(int64_t P, int64_t H, int64_t W, int flags);
int b2_shoot_bwd_ex(const b2_shoot_bwd_args* args, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- loss epilogue of the path (RegistrationReconstructionLoss, registration_losses.py:22-28) ----
 * Op-level form of b2_shoot_args.loss_terms: terms (P,2) = {sum_x (tar - sdef)^2, sum_x vel . m0} per pair from
 * dense sdef, tar (P,1,H,W) and vel, m0 (P,2,H,W).  Either pair of inputs may be NULL (its term is 0). */
int b2_recon_loss_terms(const float* sdef, const float* tar, const float* vel, const float* m0,
                        float* terms, int64_t P, int64_t H, int64_t W, void* stream);
/* Adjoint of sq[p] = sum_x (tar - interp(src, u))^2 without materialising sdef or its gradient image:
 * recomputes sdef from the taps, du (+)= g_sq[p] * 2 (sdef - tar) * d sdef/du  (accumulate != 0 adds to du),
 * dsrc (optional, zero-filled by the call; (B,1,H,W) or (P,1,H,W) with src_per_pair) receives the splat.
 * src / tar addressing as in b2_shoot_args (slice strides; 0 = dense). */
int b2_warp_sqerr_bwd(const float* g_sq, const float* src, const float* tar, const float* u,
                      float* du, float* dsrc, int64_t B, int64_t T1, int64_t H, int64_t W,
                      int src_per_pair, int64_t src_slice_stride, int64_t tar_slice_stride,
                      int background, int accumulate, void* stream);

/* ---- augmentation in front of the path (modules/data/augmentation/affine.py:24-87, rotate then translate) ----
 * vol, out (B,1,T,H,W).  xform (B,6) float64 on the device: per slice the inverse map of
 * skimage.transform.rotate(mask, -n*360/126, order=0, resize=False): (col_in,row_in) = M (col_out,row_out,1),
 * nearest neighbour, outside -> 0.  shift (B,2) int32 {translate_y, translate_x} (np.roll on rows / cols, applied
 * after the rotation) or NULL.  Index selection in float64 without contraction: bit-exact against numpy. */
int b2_augment_volume(const float* vol, float* out, const double* xform, const int32_t* shift,
                      int64_t B, int64_t T, int64_t H, int64_t W, void* stream);
/* out[b, (k + n[b]) mod R, :] = S[b, k, :]: np.roll(strain, n, axis=0) / np.roll(TOS, n) per sample
 * (affine.py:74,78); S (B,R,C), n (B) int32 on the device. */
int b2_roll_rows(const float* S, float* out, const int32_t* n, int64_t B, int64_t R, int64_t C, void* stream);

/* ---- slice regrouping behind the path (joint_registration_regression_trainer.py:54-120) ----
 * u (P,C,H,W) per-pair fields; pair_slot (P) int32 on the device: slice*F + frame position of pair p, or -1 for a
 * pair beyond the F frames kept.  out (n_slices,C,F,H,W): frames cropped to F, missing frames zero (the call
 * zero-fills out first). */
int b2_regroup_pairs(const float* u, const int32_t* pair_slot, float* out, int64_t P, int64_t n_slices,
                     int64_t F, int64_t C, int64_t H, int64_t W, void* stream);
/* Adjoint (the reference builds this tensor with differentiable stack / permute / pad and backpropagates the LMA
 * loss through it, joint_registration_regression_trainer.py:290-320): gu[p] = gout[pair_slot[p]], zero for a
 * dropped pair. */
int b2_regroup_pairs_bwd(const float* gout, const int32_t* pair_slot, float* gu, int64_t P, int64_t n_slices,
                         int64_t F, int64_t C, int64_t H, int64_t W, void* stream);

/* ---- binary masks over PCIe as one byte per pixel (host-buffer entry point) ----
 * The reference's cine inputs are fp32 volumes holding exactly 0 or 1 (README.md:21, joint_dataset.py:61-89).
 * b2_pack_binary_u8_host runs on the HOST (multi-threaded): dst[i] = (uint8_t)src[i]; returns 1 if every value was
 * exactly 0.0f or 1.0f, 0 otherwise (the caller must then copy the fp32 data), < 0 on bad arguments.
 * b2_unpack_u8 widens on the device: out[i] = (float)in[i]; n multiple of 4, in 4-byte / out 16-byte aligned. */
int b2_pack_binary_u8_host(const float* src_host, uint8_t* dst_host, int64_t n, int threads);
int b2_unpack_u8(const uint8_t* in, float* out, int64_t n, void* stream);
/* The same masks as ONE BIT per pixel (a dataset that stores numpy.packbits(mask, axis=-1): most significant bit
 * first): out[i] = (float)bit i of in; n pixels, multiple of 8, out 16-byte aligned. */
int b2_unpack_bits(const uint8_t* in, float* out, int64_t n, void* stream);

/* Device properties the host side needs for grid sizing / reporting. */
int b2_device_sm_count(int device);
/* 256x256 shooting: how many 4-CTA clusters of the fused kernel are co-resident on the current device (0 = clusters
 * unavailable) and how many SMs they leave without a CTA (a 4-SM cluster must sit inside one GPC, so GPCs whose SM
 * count is not a multiple of 4 strand SMs: 33 clusters / 16 idle SMs on a 148-SM B200).  The host side uses the idle
 * SMs for op-level work on a second stream (shooting.py, `idle_sm_split`). */
int b2_shoot_cluster_occupancy(int* clusters, int* idle_sms);

#ifdef __cplusplus
}
#endif
#endif /* B2LDDMM_H */

"""Geodesic shooting: ``EPDiff_step`` / ``expmap`` (lagomorph names) and the fused
shoot + warp + strain operator behind ``forward_volume``.

Reference call site: ``joint_register_strainmat_model.forward_volume(src_vol, tar_vol)``
(/root/reference/modules/trainer/joint_registration_strainmat_LMA.py:307); the
algorithm is SURVEY.md Appendix A.6.  The forward runs as ONE persistent CUDA
kernel (csrc/shoot.cu); the backward is the EPDiff adjoint sweep over the saved
trajectory (``b2_shoot_bwd``).
"""
from __future__ import annotations

import ctypes as C
import math

import torch
from torch.autograd.function import once_differentiable

import contextlib

from . import _lib
from ._lib import ShootArgs, ShootBwdArgs, check, lib, ptr, require_cuda, stream
from .ops import BG, Ad_star, FluidMetric, compose_disp_vel
from .strain import N_SECTORS, Frame, mask_moments

# Path selection is explicit: these are the B2_FLAG_* words handed to b2_shoot_fwd / b2_shoot_bwd_ex.  The fused
# kernels are the default wherever they exist; `force_oplevel` selects the op-level kernel sequence (path B) for
# A/B comparisons and tests.  No environment variable is read.
_flags = {"fwd": 0, "bwd": 0}
# The seeds of dL/du^S (adjoint of the strain-matrix reduction and of the squared-error term) are taken in the prologue
# of the fused adjoint kernel where that exists (square grids up to 128x128); False runs them as separate kernels
# through a gradient image (always the case at 256x256, on the op-level path, or when dL/dsrc is wanted).
fuse_seeds = True
# 256x256 inference: a 4-SM cluster has to sit inside one GPC, so the fused cluster kernel leaves SMs without a CTA
# (33 clusters, 16 idle SMs on a 148-SM B200).  With `idle_sm_split` the last slices of a batch go through the op-level
# kernel sequence on a second stream: the persistent cluster kernel is launched first and owns its SMs, so the
# op-level kernels land on exactly the stranded ones.  `idle_sm_pair_cost` = time of one op-level pair on 16 idle SMs in
# units of one cluster's time per pair (measured on B200); the split minimises the longer arm.
idle_sm_split = True
idle_sm_pair_cost = 0.23
idle_sm_pair_granular = True      # cut the batch at a pair, not at a slice boundary (_idle_split_pairs)
idle_sm_pair_cost_concurrent = 0.27   # op-level pair on 16 idle SMs WHILE the cluster kernel runs, in cluster rounds
_side_streams = {}
_cluster_occ = {}


@contextlib.contextmanager
def force_oplevel(fwd: bool = False, bwd: bool = False):
    """Within the block, run the forward and/or the adjoint as the op-level kernel sequence (B2_FLAG_OPLEVEL)."""
    old = dict(_flags)
    _flags["fwd"] = _lib.FLAG_OPLEVEL if fwd else 0
    _flags["bwd"] = _lib.FLAG_OPLEVEL if bwd else 0
    try:
        yield
    finally:
        _flags.update(old)


def EPDiff_step(metric: FluidMetric, m0, dt, phiinv, mommask=None, background="clamp"):
    """One EPDiff step; ``lagomorph.EPDiff_step(metric, m0, dt, phiinv, mommask=None)``."""
    m = Ad_star(phiinv, m0, background)
    if mommask is not None:
        m = m * mommask
    v = metric.sharp(m)
    return compose_disp_vel(phiinv, v, -dt, background)


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


def _alloc_outputs(P, B, T1, H, W, dev, want, v0_is_momentum, n_sectors, n_frames, num_steps, save_traj):
    new = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)  # noqa: E731
    out = {"u": new(P, 2, H, W)}
    if want.get("m0") and not v0_is_momentum:
        out["m0"] = new(P, 2, H, W)
    if want.get("vel"):
        out["vel"] = new(P, 2, H, W)
    if want.get("sdef"):
        out["sdef"] = new(P, 1, H, W)
    if want.get("S"):
        out["S"] = new(B, 1, n_sectors, n_frames)
        out["counts"] = torch.empty((B, n_sectors, T1), dtype=torch.int32, device=dev)
    if want.get("loss_terms"):
        out["loss_terms"] = new(P, 2)
    if save_traj:
        out["traj"] = new(num_steps, 2, P, 2, H, W)
    return out


def _idle_split_slices(B, T1, dev):
    """Trailing slices of a 256x256 inference batch that go to the op-level path on the SMs the clusters strand
    (0 = no split): minimise max(rounds of the cluster kernel, op-level time), both in units of one cluster round."""
    if not idle_sm_split or B < 2:
        return 0
    occ = _cluster_occ.get(dev.index)
    if occ is None:
        n, idle = C.c_int(0), C.c_int(0)
        check(lib().b2_shoot_cluster_occupancy(C.byref(n), C.byref(idle)), "b2_shoot_cluster_occupancy")
        occ = _cluster_occ[dev.index] = (n.value, idle.value)
    ncl, idle = occ
    if ncl < 1 or idle < 4:
        return 0
    cost = idle_sm_pair_cost * 16.0 / idle
    whole = math.ceil(B * T1 / ncl)
    best, best_t = 0, float(whole)
    for b2 in range(1, B // 2 + 1):
        t = max(math.ceil((B - b2) * T1 / ncl), b2 * T1 * cost)
        if t < best_t:
            best, best_t = b2, t
    return best if best_t <= 0.97 * whole else 0


def _idle_split_pairs(B, T1, dev, b2):
    """Refine the slice-granular split ``b2`` (:func:`_idle_split_slices`) to PAIR granularity: the number of trailing
    frame-pairs for the op-level arm.  A slice is 1.5 cluster rounds at configs[3] (49 pairs on 33 clusters), so whole
    slices leave one arm up to a round and a half behind; ``b2_shoot_args.pair_begin / pair_count`` lets the two arms
    cut a slice anywhere (strain-matrix columns are written per pair).  Rule (``tools/sweep_idle_split.py``, DESIGN.md
    section 6): the fewest cluster rounds C for which the op-level arm still ends inside the cluster arm
    (``idle_sm_pair_cost_concurrent`` rounds per pair while the cluster kernel runs), and for that C the fewest
    op-level pairs, P - 33 C."""
    occ = _cluster_occ.get(dev.index)
    if not idle_sm_pair_granular or occ is None or occ[0] < 1 or occ[1] < 4:
        return b2 * T1
    ncl, idle = occ
    cost = idle_sm_pair_cost_concurrent * 16.0 / idle
    P = B * T1
    best = 0
    for rounds in range(math.ceil(P / ncl) - 1, 0, -1):
        p2 = P - rounds * ncl
        if p2 > P // 2 or cost * p2 > rounds:
            break
        best = p2
    return best if best > 0 else b2 * T1


def _launch_shoot_split(P2, v0, src, tar, moments, frame, metric, num_steps, T, background, n_sectors, n_frames,
                        B, T1, want, src_per_pair, src_ss, tar_ss):
    """256x256 inference with the last ``P2`` frame-pairs on the op-level path (second stream, idle SMs); the outputs
    are one set of slice-major tensors, each arm writes its own pairs (``pair_begin`` / ``pair_count`` of
    ``b2_shoot_args``; the cut may fall inside a slice).  ``src`` / ``tar``: the strided cine views."""
    P, _, H, W = v0.shape
    dev = v0.device
    out = _alloc_outputs(P, B, T1, H, W, dev, want, False, n_sectors, n_frames, num_steps, False)
    P1 = P - P2
    bs = P1 // T1                                    # first slice the op-level arm touches (whole or in part)

    def rows(lo, hi):
        return {k: (t[lo:hi] if k in ("S", "counts") else t[lo * T1:hi * T1]) for k, t in out.items()}

    cur = torch.cuda.current_stream(dev)
    side = _side_streams.get(dev.index)
    if side is None:
        side = _side_streams[dev.index] = torch.cuda.Stream(dev)
    ready = cur.record_event()                       # inputs complete
    _launch_shoot(v0, src, tar, moments, frame, metric, num_steps, T, background, n_sectors, n_frames,
                  B, T1, want, False, src_per_pair, False, out=out, src_slice_stride=src_ss,
                  tar_slice_stride=tar_ss, pair_range=(0, P1))
    side.wait_event(ready)
    with torch.cuda.stream(side):
        nb = B - bs
        src2 = src[bs:].reshape(nb * T1, 1, H, W).contiguous() if src_per_pair else src[bs:].contiguous()
        tar2 = tar[bs:].reshape(nb * T1, 1, H, W).contiguous()
        mom2 = moments[bs:].contiguous() if moments is not None else None
        _launch_shoot(v0[bs * T1:], src2, tar2, mom2, frame, metric, num_steps, T, background, n_sectors, n_frames,
                      nb, T1, want, False, src_per_pair, False, out=rows(bs, B), first_slice=bs,
                      flags=_lib.FLAG_OPLEVEL, pair_range=(P1 - bs * T1, P2))
    cur.wait_stream(side)
    return out


def _launch_shoot(v0, src, tar, moments, frame, metric, num_steps, T, background, n_sectors, n_frames,
                  B, T1, want, v0_is_momentum, src_per_pair, save_traj, out=None, ws=None,
                  src_slice_stride=0, tar_slice_stride=0, first_slice=0, flags=None, pair_range=None):
    """Run ``b2_shoot_fwd``; outputs are allocated here unless ``out`` (contiguous tensors) is given.
    ``frame``: :class:`strain.Frame` of the batch (``first_slice`` = offset of this launch's slices in it) or None.
    The caller has made the tensors' device current (``_lib.device_guard`` / ``on_device``)."""
    P, _, H, W = v0.shape
    dev = v0.device
    if out is None:
        out = _alloc_outputs(P, B, T1, H, W, dev, want, v0_is_momentum, n_sectors, n_frames, num_steps, save_traj)
    a = ShootArgs()
    a.v0, a.src, a.tar = v0.data_ptr(), (src.data_ptr() if src is not None else None), \
        (tar.data_ptr() if tar is not None else None)
    a.moments = moments.data_ptr() if moments is not None else None
    if frame is not None:
        fs = frame.c_struct(first_slice)
        a.table, a.table_slice_stride, a.theta0, a.clockwise = fs.table, fs.table_slice_stride, fs.theta0, fs.clockwise
    a.flags = _flags["fwd"] if flags is None else flags
    if pair_range is not None and tuple(pair_range) != (0, P):      # (begin, count) of the batch's pairs
        a.pair_begin, a.pair_count = int(pair_range[0]), int(pair_range[1])
    for k in ("m0", "vel", "u", "sdef", "S", "counts", "traj", "loss_terms"):
        setattr(a, k, out[k].data_ptr() if k in out else None)
    a.B, a.T1, a.H, a.W = B, T1, H, W
    a.src_slice_stride, a.tar_slice_stride = int(src_slice_stride), int(tar_slice_stride)
    a.num_steps, a.src_per_pair, a.v0_is_momentum = int(num_steps), int(src_per_pair), int(v0_is_momentum)
    a.n_sectors, a.n_frames, a.background = int(n_sectors), int(n_frames), int(background)
    a.alpha, a.beta, a.gamma, a.T = metric.alpha, metric.beta, metric.gamma, float(T)
    nbytes = lib().b2_shoot_workspace_bytes_flags(B, T1, H, W, int(num_steps), a.flags)
    if nbytes <= 0:
        check(-4, "b2_shoot_workspace_bytes")
    if ws is None or ws.numel() < nbytes:
        ws = _workspace(nbytes, dev)
    check(lib().b2_shoot_fwd(C.byref(a), ptr(ws), ws.numel(), stream()), "b2_shoot_fwd")
    fused = _fused_size(H, W) and not a.flags        # one persistent kernel (256: one 4-CTA cluster per pair)
    _lib.count_launch(1 if fused else 3 + 3 * int(num_steps) + 2)   # path B: flat, 3 kernels per step, warp, strain
    return out


def _shoot_bwd(gu, gvel, gm0, m0, traj, metric, num_steps, T, background, v0_is_momentum, g_reg=None, seeds=None):
    """``seeds``: dict of the fused-seed fields of ``b2_shoot_bwd_args`` (tensors / ints) or None."""
    P, _, H, W = m0.shape
    gv0 = torch.empty_like(m0)
    a = ShootBwdArgs()
    for k, t in (("gu", gu), ("gvel", gvel), ("gm0", gm0), ("g_reg", g_reg), ("m0", m0), ("traj", traj), ("gv0", gv0)):
        setattr(a, k, t.data_ptr() if t is not None else None)
    a.P, a.H, a.W = P, H, W
    a.num_steps, a.background, a.v0_is_momentum, a.flags = int(num_steps), int(background), int(v0_is_momentum), _flags["bwd"]
    a.alpha, a.beta, a.gamma, a.T = metric.alpha, metric.beta, metric.gamma, float(T)
    for k, v in (seeds or {}).items():
        setattr(a, "seed_" + k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
    nbytes = lib().b2_shoot_bwd_workspace_bytes_flags(P, H, W, a.flags)       # sized per path (fused: resident CTAs)
    if nbytes <= 0:
        check(-4, "b2_shoot_bwd_workspace_bytes")
    ws = _workspace(nbytes, m0.device)
    check(lib().b2_shoot_bwd_ex(C.byref(a), ptr(ws), nbytes, stream()), "b2_shoot_bwd_ex")
    _lib.count_launch(1 if _fused_bwd_size(H, W) else 6 * int(num_steps) + 1)
    return gv0


class ExpmapFunction(torch.autograd.Function):
    """u = expmap(metric, m0): fused forward, adjoint sweep backward."""

    @staticmethod
    @_lib.device_guard
    def forward(ctx, m0, metric, T, num_steps, background):
        m0 = m0.contiguous()
        require_cuda(m0)
        P = m0.shape[0]
        need = ctx.needs_input_grad[0]
        out = _launch_shoot(m0, None, None, None, None, metric, num_steps, T, background, 3, 1, P, 1,
                            {}, True, False, need)
        if need:
            ctx.save_for_backward(m0, out["traj"])
        ctx.cfg = (metric, num_steps, T, background)
        return out["u"]

    @staticmethod
    @once_differentiable
    @_lib.device_guard
    def backward(ctx, gu):
        m0, traj = ctx.saved_tensors
        metric, num_steps, T, background = ctx.cfg
        g = _shoot_bwd(gu.contiguous(), None, None, m0, traj, metric, num_steps, T, background, True)
        return g, None, None, None, None


def expmap(metric: FluidMetric, m0, T=1.0, num_steps=10, phiinv=None, mommask=None, checkpoints=False,
           background="clamp"):
    """``lagomorph.expmap(metric, m0, T=1.0, num_steps=10, phiinv=None, mommask=None, checkpoints=False)``.

    Returns the inverse-map displacement u with phi^-1(x) = x + u(x).  With the
    default ``phiinv=None, mommask=None`` the whole geodesic is one fused kernel;
    otherwise it is the step-by-step composition of :func:`EPDiff_step`.
    ``checkpoints``: upstream wraps each step in ``torch.utils.checkpoint`` to trade recomputation for memory.
    The fused path already keeps the minimum the adjoint needs ((u_s, v_s) per step) and recomputes the rest
    inside the adjoint kernel, so the flag changes nothing there; on the step-by-step path it checkpoints every
    step exactly like upstream (same values, same gradient).
    """
    if phiinv is None and mommask is None:
        return ExpmapFunction.apply(m0, metric, float(T), int(num_steps), BG[background])
    u = torch.zeros_like(m0) if phiinv is None else phiinv
    dt = T / num_steps
    for _ in range(num_steps):
        if checkpoints and torch.is_grad_enabled() and (m0.requires_grad or u.requires_grad):
            from torch.utils.checkpoint import checkpoint
            u = checkpoint(lambda m, p: EPDiff_step(metric, m, dt, p, mommask=mommask, background=background),
                           m0, u, use_reentrant=False)
        else:
            u = EPDiff_step(metric, m0, dt, u, mommask=mommask, background=background)
    return u


class ShootWarpStrainFunction(torch.autograd.Function):
    """(v0, src, tar) -> (m0, vel, u, sdef, S): the body of ``forward_volume`` as one kernel.

    ``src`` / ``tar`` may be strided views of one cine volume (``src_ss`` / ``tar_ss`` = elements between
    consecutive slices, 0 = contiguous): the kernel reads the frames in place, nothing is repeated or copied.
    """

    @staticmethod
    @_lib.device_guard
    def forward(ctx, v0, src, tar, moments, frame, metric, num_steps, T, background, n_sectors, n_frames, B, T1,
                src_per_pair, with_strain, src_ss, tar_ss, with_loss=False):
        v0 = v0.contiguous()
        require_cuda(v0)
        require_cuda(src if src_ss == 0 else None, tar if tar_ss == 0 else None)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        want = {"m0": True, "vel": True, "sdef": True, "S": with_strain, "loss_terms": with_loss}
        H, W = v0.shape[-2:]
        # (a stream capture keeps the plain single-stream launch)
        b2 = _idle_split_slices(B, T1, v0.device) if (H == 256 and _fused_size(H, W) and not need and src_ss
                                                      and not torch.cuda.is_current_stream_capturing()) else 0
        if b2:
            out = _launch_shoot_split(_idle_split_pairs(B, T1, v0.device, b2), v0, src, tar, moments if with_strain else None,
                                      frame if with_strain else None, metric, num_steps, T, background, n_sectors,
                                      n_frames, B, T1, want, src_per_pair, src_ss, tar_ss)
        else:
            out = _launch_shoot(v0, src, tar, moments if with_strain else None, frame if with_strain else None, metric,
                                num_steps, T, background, n_sectors, n_frames, B, T1, want, False,
                                src_per_pair, need, src_slice_stride=src_ss, tar_slice_stride=tar_ss)
        ctx.cfg = (metric, num_steps, T, background, n_sectors, n_frames, B, T1, src_per_pair, with_strain,
                   src_ss, tar_ss)
        ctx.set_materialize_grads(False)      # unused outputs arrive as None in backward, not as zero tensors
        if need:
            ctx.save_for_backward(out["m0"], out["u"], out["traj"], src, tar, moments,
                                  out.get("counts", torch.empty(0, device=v0.device)))
            ctx.frame = frame
        if with_strain:
            S = out["S"]
        else:
            S = torch.zeros((B, 1, n_sectors, n_frames), device=v0.device)
            ctx.mark_non_differentiable(S)
        if with_loss:
            lt = out["loss_terms"]
        else:
            lt = torch.zeros((0, 2), device=v0.device)
            ctx.mark_non_differentiable(lt)
        return out["m0"], out["vel"], out["u"], out["sdef"], S, lt

    @staticmethod
    @once_differentiable
    @_lib.device_guard
    def backward(ctx, gm0, gvel, gu, gsdef, gS, glt=None):
        m0, u, traj, src, tar, moments, counts = ctx.saved_tensors
        metric, num_steps, T, background, n_sectors, n_frames, B, T1, src_per_pair, with_strain, src_ss, tar_ss = ctx.cfg
        P, _, H, W = m0.shape
        want_dsrc = ctx.needs_input_grad[1]
        gu_tot = None          # dL/du^S, accumulated in place by the kernels where they support it
        dsrc = None

        def add(acc, x):
            return x if acc is None else acc.add_(x)

        # seeds of dL/du^S inside the adjoint kernel (no gradient image, no extra launches) where it has the prologue
        seeds = None
        fused_seeds = (fuse_seeds and ctx.needs_input_grad[0] and not want_dsrc and _fused_bwd_size(H, W) and H <= 128
                       and ((with_strain and gS is not None) or glt is not None))
        keep = []              # tensors that must outlive the launch
        if fused_seeds:
            seeds = {"u": u, "tar": tar, "T1": T1, "tar_slice_stride": int(tar_ss)}
            if with_strain and gS is not None:
                gS_c = gS.contiguous()
                keep.append(gS_c)
                fs = ctx.frame.c_struct()
                seeds.update({"gS": gS_c, "counts": counts, "moments": moments, "table": fs.table,
                              "table_slice_stride": fs.table_slice_stride, "theta0": fs.theta0, "clockwise": fs.clockwise,
                              "n_sectors": n_sectors, "n_frames": n_frames})
        elif with_strain and gS is not None:
            du = torch.empty_like(u)
            tar_c = tar.reshape(B, T1, H, W).contiguous()
            gS_c = gS.contiguous()          # named: a temporary would be freed before the launch is enqueued
            fs = ctx.frame.c_struct()
            check(lib().b2_strain_sector_bwd_ex(ptr(gS_c), ptr(u), ptr(tar_c), ptr(moments), C.byref(fs),
                                                ptr(counts), ptr(du), B, T1, H, W, n_sectors, n_frames, stream()),
                  "b2_strain_sector_bwd_ex")
            _lib.count_launch()
            gu_tot = du
        g_reg = None
        if glt is not None:
            # loss epilogue: adjoint of sum (tar - interp(src,u))^2 straight from (src, tar, u) - strided volumes are
            # read in place - and the closed-form regularisation gradient inside the adjoint kernel (g_reg)
            g_sq = glt[:, 0].contiguous()
            g_reg = glt[:, 1].contiguous()
            if fused_seeds:
                keep.append(g_sq)
                seeds.update({"g_sq": g_sq, "src": src, "src_slice_stride": int(src_ss), "src_per_pair": int(src_per_pair)})
            else:
                acc = gu_tot is not None
                if not acc:
                    gu_tot = torch.empty_like(u)
                d2 = torch.empty((P if src_per_pair else B, 1, H, W), device=u.device) if want_dsrc else None
                check(lib().b2_warp_sqerr_bwd(ptr(g_sq), ptr(src), ptr(tar), ptr(u), ptr(gu_tot), ptr(d2), B, T1, H, W,
                                              int(src_per_pair), int(src_ss), int(tar_ss), background, int(acc), stream()),
                      "b2_warp_sqerr_bwd")
                _lib.count_launch()
                dsrc = add(dsrc, d2) if d2 is not None else dsrc
        if gsdef is not None:
            # op-level adjoint kernels take dense batches
            src_c = src.reshape(P, 1, H, W).contiguous() if src_per_pair else src.contiguous()
            du = torch.empty_like(u)
            d1 = torch.empty_like(src_c) if want_dsrc else None
            gsdef_c = gsdef.contiguous()
            if src_per_pair:
                check(lib().b2_interp_bwd(ptr(gsdef_c), ptr(src_c), ptr(u), ptr(d1), ptr(du), P, P, P, 1,
                                          H, W, 1.0, background, stream()), "b2_interp_bwd")
            else:
                check(lib().b2_warp_bwd(ptr(gsdef_c), ptr(src_c), ptr(u), ptr(d1), ptr(du), B, T1, 1,
                                        H, W, 1.0, background, stream()), "b2_warp_bwd")
            _lib.count_launch()
            gu_tot = add(gu_tot, du)
            dsrc = add(dsrc, d1) if d1 is not None else dsrc
        if gu is not None:
            gu_tot = gu.contiguous() if gu_tot is None else gu_tot.add_(gu)
        gv0 = None
        if ctx.needs_input_grad[0]:
            gv0 = _shoot_bwd(gu_tot, gvel.contiguous() if gvel is not None else None,
                             gm0.contiguous() if gm0 is not None else None, m0, traj, metric, num_steps, T,
                             background, False, g_reg=g_reg, seeds=seeds)
        if dsrc is not None:
            dsrc = dsrc.reshape(src.shape)
        return (gv0, dsrc) + (None,) * 16


def _fused_size(H, W):
    """Sizes served by a persistent fused forward kernel (they read strided cine volumes in place): one CTA per pair
    up to 128x128, one 4-CTA cluster per pair at 256x256 - unless ``force_oplevel(fwd=True)`` is active."""
    return H == W and H in (16, 32, 64, 128, 256) and not _flags["fwd"]


def _fused_bwd_size(H, W):
    return H == W and H in (16, 32, 64, 128, 256) and not _flags["bwd"]


def _rows_dense(t):
    return t.dtype == torch.float32 and t.is_cuda and t.stride(-1) == 1 and t.stride(-2) == t.shape[-1]


def shoot_warp_strain(v0, src_vol, tar_vol, metric: FluidMetric, num_steps=10, T=1.0, n_sectors=N_SECTORS,
                      n_frames=40, background="clamp", with_strain=True, loss_terms=False, theta0=None,
                      clockwise=None):
    """Fused hot path for a batch of slices.

    v0: (B*T1, 2, H, W) initial velocities, slice-major; src_vol, tar_vol: (B,1,T1,H,W)
    (the outputs of ``split_vol_to_registration_pairs(..., 'Lagrangian', output_dim=3)``;
    only frame 0 of ``src_vol`` is read - the repeat is never materialised - and ``tar_vol`` may be the
    strided view ``vol[:, :, 1:]``: the kernel reads the cine volume in place).
    Returns the dict ``forward_volume`` hands to the trainer plus 'displacement'.  With ``loss_terms=True`` the
    dict also holds 'registration_loss_terms' (P,2) = per pair {sum (tar - Sdef)^2, sum v.m}: the two reductions of
    RegistrationReconstructionLoss taken inside the shooting kernel (see :mod:`losses`); their backward needs no
    seed tensors.  ``theta0`` (radians) / ``clockwise``: per-slice sector frame of the strain matrix rows
    (:mod:`strain`; DENSE_utils.py:196-204 of the reference) - scalars or B entries; default 0 / clockwise.
    """
    B, Cc, T1, H, W = tar_vol.shape
    if Cc != 1 or v0.shape != (B * T1, 2, H, W):
        raise _lib.B2Error(f"shape mismatch: v0 {tuple(v0.shape)}, tar_vol {tuple(tar_vol.shape)}")
    if tuple(src_vol.shape) != (B, Cc, T1, H, W):
        raise _lib.B2Error(f"shape mismatch: src_vol {tuple(src_vol.shape)}, tar_vol {tuple(tar_vol.shape)}")
    mask0 = src_vol[:, 0, 0]                                   # frame-0 mask: centroid + sectors of the slice
    # Lagrangian split (modules/data/__init__.py:108-110): every pair of a slice shares frame 0 - an expanded
    # view has frame stride 0.  Otherwise (Eulerian split, or a materialised repeat) each pair has its own source.
    shared = T1 == 1 or src_vol.stride(2) == 0
    src = src_vol[:, :, 0] if shared else src_vol
    src_ss = tar_ss = 0
    strided_ok = (_fused_size(H, W) and _rows_dense(src) and _rows_dense(tar_vol) and tar_vol.stride(2) == H * W
                  and (shared or src_vol.stride(2) == H * W))
    if strided_ok:
        src_ss = src.stride(0) if B > 1 else T1 * H * W
        tar_ss = tar_vol.stride(0) if B > 1 else T1 * H * W
        tar = tar_vol
    else:
        src = src.contiguous() if shared else src_vol.reshape(B * T1, 1, H, W).contiguous()
        tar = tar_vol.reshape(B * T1, 1, H, W).contiguous()
    moments = mask_moments(mask0.contiguous()) if with_strain else None
    frame = Frame(n_sectors, B, v0.device, theta0, clockwise) if with_strain else None
    m0, vel, u, sdef, S, lt = ShootWarpStrainFunction.apply(
        v0, src, tar, moments, frame, metric, int(num_steps), float(T), BG[background], int(n_sectors),
        int(n_frames), B, T1, not shared, bool(with_strain), int(src_ss), int(tar_ss), bool(loss_terms))
    out = {
        "strain_matrix": S,
        "deformed_source": sdef.reshape(B, 1, T1, H, W),
        "velocity": vel,
        "momentum": m0,
        "displacement": u,
    }
    if loss_terms:
        out["registration_loss_terms"] = lt
    return out


def shoot_warp_pairs(v0, src, tar, metric: FluidMetric, num_steps=10, T=1.0, background="clamp", loss_terms=False):
    """Pairwise contract (/root/reference/modules/trainer/reg_trainer.py:45,222-225): src, tar (P,1,H,W)."""
    P, _, H, W = v0.shape
    m0, vel, u, sdef, _, lt = ShootWarpStrainFunction.apply(
        v0, src.contiguous(), tar.contiguous(), None, None, metric, int(num_steps), float(T), BG[background], 3, 1,
        P, 1, True, False, 0, 0, bool(loss_terms))
    out = {"displacement": u, "velocity": vel, "momentum": m0, "deformed_source": sdef}
    if loss_terms:
        out["registration_loss_terms"] = lt
    return out


class PipelineResult:
    """Handle of one :meth:`HostPipeline.submit`: ``get()`` blocks until the strain matrices of that call are
    visible on the host and returns them (pinned tensor owned by the pipeline, valid until the call after next)."""

    def __init__(self, S_host, done):
        self._S, self.done = S_host, done

    def get(self):
        self.done.synchronize()
        return self._S


class HostPipeline:
    """Host-buffer entry point of the hot path: pinned host inputs in, strain matrices on the host out.

    The batch is cut into chunks of whole slices; the H2D copy of chunk i+1 (copy stream) overlaps the
    fused shooting kernel of chunk i (compute stream), with triple-buffered device staging, so a step
    costs max(PCIe time, kernel time) instead of their sum.  Device outputs of the whole batch stay
    available in ``self.out`` (same keys as :func:`shoot_warp_strain`).  Inference only (no autograd).

    ``pipe(v0_host, vol_host)`` returns the (B,1,n_sectors,n_frames) strain matrices on the host AFTER the
    device-to-host copy has completed (it waits on an event recorded behind the copy).  ``pipe.submit(...)`` is
    the streaming form: it returns a :class:`PipelineResult` at once so that the next call's uploads overlap this
    call's kernels; the inputs must stay untouched until ``result.get()`` returns, and two result buffers rotate,
    so at most two calls may be outstanding.

    Masks.  The reference's cine inputs are binary myocardium masks (README.md:21): ``vol_host`` may be ``uint8`` /
    ``bool`` (one byte per pixel over PCIe, widened on the device by ``b2_unpack_u8`` - no host pass at all) or
    ``float32``.  For fp32 volumes ``pack_masks=True`` narrows each chunk on the host (multi-threaded, verified to
    be exactly 0/1, otherwise that chunk is copied as fp32) while the ``v0`` copy of the same chunk occupies the bus;
    whether that pays is MEASURED: the pass is timed against the copy it has to hide behind and switched off for
    later calls when it does not fit (few or busy host cores).  A dataset that stores its masks as ONE BIT per pixel
    hands over ``numpy.packbits(mask, axis=-1)``: ``uint8`` of shape (B,1,T,H,W/8), most significant bit first
    (W a multiple of 8), widened on the device by ``b2_unpack_bits`` - no host pass, an eighth of the bytes of the
    ``uint8`` route.  Results are bit-identical on every route.
    ``self.h2d_bytes`` is the number of bytes the last call copied to the device.
    """

    def __init__(self, B, T, H, W, metric: FluidMetric, num_steps=10, T_end=1.0, n_sectors=N_SECTORS, n_frames=40,
                 chunk_slices=None, device=None, background="clamp", pack_masks=True, pack_threads=0, theta0=None,
                 clockwise=None, n_stages=3):
        self.dev = torch.device(device if device is not None else torch.cuda.current_device())
        if self.dev.index is None:
            self.dev = torch.device("cuda", torch.cuda.current_device())
        self.B, self.T, self.T1, self.H, self.W = B, T, T - 1, H, W
        self.metric, self.num_steps, self.T_end = metric, int(num_steps), float(T_end)
        self.n_sectors, self.n_frames, self.bg = int(n_sectors), int(n_frames), BG[background]
        if chunk_slices is None:      # four equal chunks: measured best at configs[1] (16 slices; 18 / 32 / 37 were slower)
            chunk_slices = -(-B // 4)
        self.chunk = max(1, min(int(chunk_slices), B))
        dev, T1, cs = self.dev, self.T1, self.chunk
        self.byte_masks_ok = (T * H * W) % 4 == 0            # b2_unpack_u8 works on whole 4-byte words
        self.pack_masks = bool(pack_masks) and self.byte_masks_ok
        if int(pack_threads) <= 0:    # CPUs this process may use, shared with the other ranks of the node
            import os
            try:
                ncpu = len(os.sched_getaffinity(0))
            except (AttributeError, OSError):
                ncpu = os.cpu_count() or 1
            local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
            pack_threads = max(1, min(16, ncpu // local_world))
        self.pack_threads = int(pack_threads)
        self._pack_s = self._pack_budget_s = 0.0
        self._pack_chunks = 0
        self.h2d_bytes = 0
        with torch.cuda.device(dev):
            self.copy_stream = torch.cuda.Stream(dev)
            # Three staging buffers: with two, the copy of chunk i+1 has to wait for the kernel of chunk i-1, which
            # ends just about when the copy of chunk i does (kernel and copy times per chunk are nearly equal) - any
            # jitter stalls the copy engine.
            self.n_stages = max(2, int(n_stages))
            self._next_stage = 0
            self.stage = [{"vol": torch.empty((cs, 1, T, H, W), device=dev),
                           "v0": torch.empty((cs * T1, 2, H, W), device=dev),
                           "ready": torch.cuda.Event(), "free": torch.cuda.Event(), "used": False}
                          for _ in range(self.n_stages)]
            if self.byte_masks_ok:
                for st in self.stage:
                    st["vol_u8"] = torch.empty(cs * T * H * W, dtype=torch.uint8, device=dev)
                    st["pack_host"] = None        # pinned scratch of the fp32 narrowing pass, allocated on first use
            P = B * T1
            self.out = _alloc_outputs(P, B, T1, H, W, dev, {"m0": True, "vel": True, "sdef": True, "S": True}, False,
                                      self.n_sectors, self.n_frames, self.num_steps, False)
            self._S_host = [torch.empty((B, 1, self.n_sectors, self.n_frames), dtype=torch.float32).pin_memory()
                            for _ in range(2)]
            self._done = [torch.cuda.Event(), torch.cuda.Event()]
            self._next_result = 0
            self.frame = Frame(self.n_sectors, B, dev, theta0, clockwise)
            nbytes = lib().b2_shoot_workspace_bytes_flags(cs, T1, H, W, self.num_steps, _flags["fwd"])
            self.ws = _workspace(nbytes, dev)

    def __call__(self, v0_host: torch.Tensor, vol_host: torch.Tensor):
        """v0_host (B*(T-1),2,H,W) pinned fp32, vol_host (B,1,T,H,W) pinned fp32 / uint8 / bool host tensors.
        Returns S on the host, complete (the device-to-host copy has finished)."""
        return self.submit(v0_host, vol_host).get()

    def submit(self, v0_host: torch.Tensor, vol_host: torch.Tensor) -> PipelineResult:
        B, T, T1, H, W, cs = self.B, self.T, self.T1, self.H, self.W, self.chunk
        bit_masks = (vol_host.dtype == torch.uint8 and W % 8 == 0 and W > 8
                     and tuple(vol_host.shape) == (B, 1, T, H, W // 8))
        if not (bit_masks or tuple(vol_host.shape) == (B, 1, T, H, W)) or tuple(v0_host.shape) != (B * T1, 2, H, W):
            raise _lib.B2Error(f"shape mismatch: v0 {tuple(v0_host.shape)}, vol {tuple(vol_host.shape)}")
        if vol_host.dtype == torch.bool:
            vol_host = vol_host.view(torch.uint8)
        byte_masks = vol_host.dtype == torch.uint8 and not bit_masks
        if not (v0_host.is_contiguous() and vol_host.is_contiguous() and v0_host.dtype == torch.float32
                and (byte_masks or bit_masks or vol_host.dtype == torch.float32)) or v0_host.is_cuda or vol_host.is_cuda:
            raise _lib.B2Error("HostPipeline expects contiguous host tensors: fp32 v0, fp32 / uint8 / bool / "
                               "bit-packed masks")
        if (byte_masks or bit_masks) and not self.byte_masks_ok:
            raise _lib.B2Error("uint8 masks need T*H*W to be a multiple of 4")
        self.h2d_bytes = 0
        with torch.cuda.device(self.dev), torch.no_grad():
            main = torch.cuda.current_stream(self.dev)
            ri = self._next_result
            self._next_result ^= 1
            S_host, done = self._S_host[ri], self._done[ri]
            for i, b0 in enumerate(range(0, B, cs)):
                b1 = min(b0 + cs, B)
                nb = b1 - b0
                st = self.stage[self._next_stage]                        # rotates across calls as well
                self._next_stage = (self._next_stage + 1) % self.n_stages
                with torch.cuda.stream(self.copy_stream):
                    if st["used"]:                                       # last kernel that read this stage is done
                        self.copy_stream.wait_event(st["free"])          # (also across consecutive calls)
                    st["v0"][: nb * T1].copy_(v0_host[b0 * T1: b1 * T1], non_blocking=True)
                    self.h2d_bytes += nb * T1 * 2 * H * W * 4
                    n = nb * T * H * W
                    # the copy stream carries copies only (back-to-back DMA); narrow masks are widened on the compute
                    # stream in front of the kernel that reads them
                    src_u8, widen = None, None
                    if bit_masks:
                        st["vol_u8"][: n // 8].copy_(vol_host[b0:b1].reshape(-1), non_blocking=True)
                        widen = lib().b2_unpack_bits
                        self.h2d_bytes += n // 8
                    elif byte_masks:
                        src_u8 = vol_host[b0:b1].reshape(-1)
                    elif self.pack_masks:
                        src_u8 = self._narrow_on_host(st, vol_host[b0:b1], n, nb * T1 * 2 * H * W * 4)
                    if bit_masks:
                        pass
                    elif src_u8 is not None:
                        st["vol_u8"][:n].copy_(src_u8[:n], non_blocking=True)
                        widen = lib().b2_unpack_u8
                        self.h2d_bytes += n
                    else:
                        st["vol"][:nb].copy_(vol_host[b0:b1], non_blocking=True)
                        self.h2d_bytes += n * 4
                    st["ready"].record(self.copy_stream)
                main.wait_event(st["ready"])
                if widen is not None:
                    check(widen(ptr(st["vol_u8"]), ptr(st["vol"]), n, C.c_void_p(main.cuda_stream)), "b2_unpack")
                    _lib.count_launch()
                vol = st["vol"][:nb]                                     # (nb,1,T,H,W): read in place by the kernel
                mom = mask_moments(vol[:, 0, 0].contiguous())
                sl = slice(b0 * T1, b1 * T1)
                out = {"u": self.out["u"][sl], "m0": self.out["m0"][sl], "vel": self.out["vel"][sl],
                       "sdef": self.out["sdef"][sl], "S": self.out["S"][b0:b1], "counts": self.out["counts"][b0:b1]}
                _launch_shoot(st["v0"][: nb * T1], vol, vol.view(-1)[H * W:], mom, self.frame, self.metric,
                              self.num_steps, self.T_end, self.bg, self.n_sectors, self.n_frames, nb, T1,
                              {}, False, False, False, out=out, ws=self.ws,
                              src_slice_stride=T * H * W, tar_slice_stride=T * H * W, first_slice=b0)
                st["free"].record(main)
                st["used"] = True
                S_host[b0:b1].copy_(self.out["S"][b0:b1], non_blocking=True)
            done.record(main)                                            # behind the last device-to-host copy
        return PipelineResult(S_host, done)

    def _narrow_on_host(self, st, vol_chunk, n, v0_bytes):
        """fp32 -> u8 on the host for one chunk; returns the pinned byte buffer, or None (copy as fp32)."""
        import time
        if st["pack_host"] is None:
            st["pack_host"] = torch.empty(self.chunk * self.T * self.H * self.W, dtype=torch.uint8).pin_memory()
        if st["used"]:
            st["ready"].synchronize()        # the copy out of this pinned pack buffer has finished
        # CPU pass (GIL released inside the call) while the v0 copy of this chunk occupies the bus
        t_pack = time.perf_counter()
        rc = lib().b2_pack_binary_u8_host(C.c_void_p(vol_chunk.data_ptr()), C.c_void_p(st["pack_host"].data_ptr()), n,
                                          self.pack_threads)
        if rc < 0:
            check(rc, "b2_pack_binary_u8_host")
        # measured decision: the pass only pays while it hides behind the v0 copy of the same chunk (~50 GB/s over
        # PCIe); a starved host (few or busy cores, many ranks per node) switches it off for later calls
        self._pack_s += time.perf_counter() - t_pack
        self._pack_budget_s += 0.8 * v0_bytes / 50e9
        self._pack_chunks += 1
        if self._pack_chunks >= 8 and self._pack_s > self._pack_budget_s:
            self.pack_masks = False
        return st["pack_host"] if rc == 1 else None

    def result(self):
        """Device outputs of the last call, keyed like :func:`shoot_warp_strain`."""
        B, T1, H, W = self.B, self.T1, self.H, self.W
        return {"strain_matrix": self.out["S"], "deformed_source": self.out["sdef"].view(B, 1, T1, H, W),
                "velocity": self.out["vel"], "momentum": self.out["m0"], "displacement": self.out["u"]}

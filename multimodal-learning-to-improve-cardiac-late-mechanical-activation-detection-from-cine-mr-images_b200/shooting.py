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

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import ShootArgs, check, lib, ptr, require_cuda, stream
from .ops import BG, Ad_star, FluidMetric, compose_disp_vel
from .strain import N_SECTORS, mask_moments, sector_table


def EPDiff_step(metric: FluidMetric, m0, dt, phiinv, mommask=None, background="clamp"):
    """One EPDiff step; ``lagomorph.EPDiff_step(metric, m0, dt, phiinv, mommask=None)``."""
    m = Ad_star(phiinv, m0, background)
    if mommask is not None:
        m = m * mommask
    v = metric.sharp(m)
    return compose_disp_vel(phiinv, v, -dt, background)


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


def _launch_shoot(v0, src, tar, moments, table, metric, num_steps, T, background, n_sectors, n_frames,
                  B, T1, want, v0_is_momentum, src_per_pair, save_traj):
    """Allocate outputs and run ``b2_shoot_fwd``.  ``want`` selects optional outputs."""
    P, _, H, W = v0.shape
    dev = v0.device
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
    if save_traj:
        out["traj"] = new(num_steps, 2, P, 2, H, W)
    a = ShootArgs()
    a.v0, a.src, a.tar = v0.data_ptr(), (src.data_ptr() if src is not None else None), \
        (tar.data_ptr() if tar is not None else None)
    a.moments = moments.data_ptr() if moments is not None else None
    a.table = table.data_ptr() if table is not None else None
    for k in ("m0", "vel", "u", "sdef", "S", "counts", "traj"):
        setattr(a, k, out[k].data_ptr() if k in out else None)
    a.B, a.T1, a.H, a.W = B, T1, H, W
    a.num_steps, a.src_per_pair, a.v0_is_momentum = int(num_steps), int(src_per_pair), int(v0_is_momentum)
    a.n_sectors, a.n_frames, a.background = int(n_sectors), int(n_frames), int(background)
    a.alpha, a.beta, a.gamma, a.T = metric.alpha, metric.beta, metric.gamma, float(T)
    nbytes = lib().b2_shoot_workspace_bytes(B, T1, H, W, int(num_steps))
    if nbytes <= 0:
        check(-4, "b2_shoot_workspace_bytes")
    ws = _workspace(nbytes, dev)
    check(lib().b2_shoot_fwd(C.byref(a), ptr(ws), nbytes, stream()), "b2_shoot_fwd")
    fused = (H == W and H in (16, 32, 64, 128))
    _lib.count_launch(1 if fused else 3 + 5 * int(num_steps) + 2)
    return out


def _shoot_bwd(gu, gvel, gm0, m0, traj, metric, num_steps, T, background, v0_is_momentum):
    P, _, H, W = m0.shape
    gv0 = torch.empty_like(m0)
    nbytes = lib().b2_shoot_bwd_workspace_bytes(P, H, W)
    ws = _workspace(nbytes, m0.device)
    check(lib().b2_shoot_bwd(ptr(gu), ptr(gvel), ptr(gm0), ptr(m0), ptr(traj), ptr(gv0), P, H, W, int(num_steps),
                             metric.alpha, metric.beta, metric.gamma, float(T), int(background),
                             int(v0_is_momentum), ptr(ws), nbytes, stream()), "b2_shoot_bwd")
    _lib.count_launch(6 * int(num_steps) + 1)
    return gv0


class ExpmapFunction(torch.autograd.Function):
    """u = expmap(metric, m0): fused forward, adjoint sweep backward."""

    @staticmethod
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
    ``checkpoints`` is accepted for signature compatibility (the fused path stores
    (u_s, v_s) per step, which is what the adjoint needs).
    """
    if phiinv is None and mommask is None:
        return ExpmapFunction.apply(m0, metric, float(T), int(num_steps), BG[background])
    u = torch.zeros_like(m0) if phiinv is None else phiinv
    dt = T / num_steps
    for _ in range(num_steps):
        u = EPDiff_step(metric, m0, dt, u, mommask=mommask, background=background)
    return u


class ShootWarpStrainFunction(torch.autograd.Function):
    """(v0, src, tar) -> (m0, vel, u, sdef, S): the body of ``forward_volume`` as one kernel."""

    @staticmethod
    def forward(ctx, v0, src, tar, moments, table, metric, num_steps, T, background, n_sectors, n_frames, B, T1,
                src_per_pair, with_strain):
        v0 = v0.contiguous()
        src = src.contiguous()
        tar = tar.contiguous()
        require_cuda(v0, src, tar)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        out = _launch_shoot(v0, src, tar, moments if with_strain else None, table if with_strain else None, metric,
                            num_steps, T, background, n_sectors, n_frames, B, T1,
                            {"m0": True, "vel": True, "sdef": True, "S": with_strain}, False, src_per_pair, need)
        ctx.cfg = (metric, num_steps, T, background, n_sectors, n_frames, B, T1, src_per_pair, with_strain)
        if need:
            ctx.save_for_backward(out["m0"], out["u"], out["traj"], src, tar, moments, table,
                                  out.get("counts", torch.empty(0, device=v0.device)))
        if with_strain:
            S = out["S"]
        else:
            S = torch.zeros((B, 1, n_sectors, n_frames), device=v0.device)
            ctx.mark_non_differentiable(S)
        return out["m0"], out["vel"], out["u"], out["sdef"], S

    @staticmethod
    @once_differentiable
    def backward(ctx, gm0, gvel, gu, gsdef, gS):
        m0, u, traj, src, tar, moments, table, counts = ctx.saved_tensors
        metric, num_steps, T, background, n_sectors, n_frames, B, T1, src_per_pair, with_strain = ctx.cfg
        P, _, H, W = m0.shape
        gu_tot = gu.contiguous().clone() if gu is not None else torch.zeros_like(u)
        dsrc = None
        if gsdef is not None:
            du = torch.empty_like(u)
            want_dsrc = ctx.needs_input_grad[1]
            dsrc = torch.empty_like(src) if want_dsrc else None
            if src_per_pair:
                check(lib().b2_interp_bwd(ptr(gsdef.contiguous()), ptr(src), ptr(u), ptr(dsrc), ptr(du), P, P, P, 1,
                                          H, W, 1.0, background, stream()), "b2_interp_bwd")
            else:
                check(lib().b2_warp_bwd(ptr(gsdef.contiguous()), ptr(src), ptr(u), ptr(dsrc), ptr(du), B, T1, 1, H, W,
                                        1.0, background, stream()), "b2_warp_bwd")
            _lib.count_launch()
            gu_tot += du
        if with_strain and gS is not None:
            du = torch.empty_like(u)
            check(lib().b2_strain_sector_bwd(ptr(gS.contiguous()), ptr(u), ptr(tar), ptr(moments), ptr(table),
                                             ptr(counts), ptr(du), B, T1, H, W, n_sectors, n_frames, stream()),
                  "b2_strain_sector_bwd")
            _lib.count_launch()
            gu_tot += du
        gv0 = None
        if ctx.needs_input_grad[0]:
            gv0 = _shoot_bwd(gu_tot, gvel.contiguous() if gvel is not None else None,
                             gm0.contiguous() if gm0 is not None else None, m0, traj, metric, num_steps, T,
                             background, False)
        return (gv0, dsrc) + (None,) * 13


def shoot_warp_strain(v0, src_vol, tar_vol, metric: FluidMetric, num_steps=10, T=1.0, n_sectors=N_SECTORS,
                      n_frames=40, background="clamp", with_strain=True):
    """Fused hot path for a batch of slices.

    v0: (B*T1, 2, H, W) initial velocities, slice-major; src_vol, tar_vol: (B,1,T1,H,W)
    (the outputs of ``split_vol_to_registration_pairs(..., 'Lagrangian', output_dim=3)``;
    only frame 0 of ``src_vol`` is read - the repeat is never materialised).
    Returns the dict ``forward_volume`` hands to the trainer plus 'displacement'.
    """
    B, Cc, T1, H, W = tar_vol.shape
    if Cc != 1 or v0.shape != (B * T1, 2, H, W):
        raise _lib.B2Error(f"shape mismatch: v0 {tuple(v0.shape)}, tar_vol {tuple(tar_vol.shape)}")
    src = src_vol[:, :, 0].contiguous()                        # (B,1,H,W) frame-0 mask
    tar = tar_vol.reshape(B * T1, 1, H, W)
    moments = mask_moments(src[:, 0]) if with_strain else None
    table = sector_table(n_sectors, v0.device) if with_strain else None
    m0, vel, u, sdef, S = ShootWarpStrainFunction.apply(
        v0, src, tar, moments, table, metric, int(num_steps), float(T), BG[background], int(n_sectors),
        int(n_frames), B, T1, False, bool(with_strain))
    return {
        "strain_matrix": S,
        "deformed_source": sdef.reshape(B, 1, T1, H, W),
        "velocity": vel,
        "momentum": m0,
        "displacement": u,
    }


def shoot_warp_pairs(v0, src, tar, metric: FluidMetric, num_steps=10, T=1.0, background="clamp"):
    """Pairwise contract (/root/reference/modules/trainer/reg_trainer.py:45,222-225): src, tar (P,1,H,W)."""
    P, _, H, W = v0.shape
    m0, vel, u, sdef, _ = ShootWarpStrainFunction.apply(
        v0, src, tar, None, None, metric, int(num_steps), float(T), BG[background], 3, 1, P, 1, True, False)
    return {"displacement": u, "velocity": vel, "momentum": m0, "deformed_source": sdef}

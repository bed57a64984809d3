"""lagomorph-compatible differentiable operators on top of the C-ABI library.

Same names, argument meaning and error behaviour as the third-party
``lagomorph`` package the reference imports
(/root/reference/modules/trainer/joint_registration_strainmat_LMA.py:5): every op
is a ``torch.autograd.Function`` on contiguous same-device fp32 CUDA tensors of
layout (N, C, H, W); bad arguments raise ``RuntimeError``.
"""
from __future__ import annotations

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import check, lib, ptr, require_cuda, stream

BG = {"clamp": 0, "zero": 1}


def _bcast_dims(I, u):
    PI, Pu = I.shape[0], u.shape[0]
    P = max(PI, Pu)
    if PI not in (1, P) or Pu not in (1, P):
        raise _lib.B2Error(f"batch sizes {PI} and {Pu} do not broadcast")
    if I.shape[-2:] != u.shape[-2:] or u.shape[1] != 2 or I.dim() != 4 or u.dim() != 4:
        raise _lib.B2Error(f"shape mismatch: I {tuple(I.shape)}, u {tuple(u.shape)}")
    return P, PI, Pu


class InterpFunction(torch.autograd.Function):
    @staticmethod
    @_lib.device_guard
    def forward(ctx, I, u, dt, background):
        I = I.contiguous()
        u = u.contiguous()
        require_cuda(I, u)
        P, PI, Pu = _bcast_dims(I, u)
        C, H, W = I.shape[1:]
        out = torch.empty((P, C, H, W), dtype=I.dtype, device=I.device)
        check(lib().b2_interp_fwd(ptr(I), ptr(u), ptr(out), P, PI, Pu, C, H, W, float(dt), background, stream()),
              "b2_interp_fwd")
        _lib.count_launch()
        ctx.save_for_backward(I, u)
        ctx.dt, ctx.bg = float(dt), background
        return out

    @staticmethod
    @once_differentiable
    @_lib.device_guard
    def backward(ctx, gout):
        I, u = ctx.saved_tensors
        gout = gout.contiguous()
        P, PI, Pu = _bcast_dims(I, u)
        C, H, W = I.shape[1:]
        dI = torch.empty_like(I) if ctx.needs_input_grad[0] else None
        du = torch.empty_like(u) if ctx.needs_input_grad[1] else None
        check(lib().b2_interp_bwd(ptr(gout), ptr(I), ptr(u), ptr(dI), ptr(du), P, PI, Pu, C, H, W, ctx.dt, ctx.bg,
                                  stream()), "b2_interp_bwd")
        _lib.count_launch()
        return dI, du, None, None


def interp(I, u, dt=1.0, background="clamp"):
    """out(x) = I(x + dt*u(x)), bilinear; ``lagomorph.interp(I, u, dt=1.0)``."""
    return InterpFunction.apply(I, u, dt, BG[background])


class SplatFunction(torch.autograd.Function):
    @staticmethod
    @_lib.device_guard
    def forward(ctx, J, u, dt, background, need_weights):
        J = J.contiguous()
        u = u.contiguous()
        require_cuda(J, u)
        P, PJ, Pu = _bcast_dims(J, u)
        C, H, W = J.shape[1:]
        out = torch.empty((P, C, H, W), dtype=J.dtype, device=J.device)
        wout = torch.empty((P, 1, H, W), dtype=J.dtype, device=J.device) if need_weights else None
        check(lib().b2_splat_fwd(ptr(J), ptr(u), ptr(out), ptr(wout), P, PJ, Pu, C, H, W, float(dt), background,
                                 stream()), "b2_splat_fwd")
        _lib.count_launch()
        ctx.save_for_backward(J, u)
        ctx.dt, ctx.bg = float(dt), background
        if need_weights:
            ctx.mark_non_differentiable(wout)
            return out, wout
        return out

    @staticmethod
    @once_differentiable
    @_lib.device_guard
    def backward(ctx, gout, gw=None):
        J, u = ctx.saved_tensors
        gout = gout.contiguous()
        P, PJ, Pu = _bcast_dims(J, u)
        C, H, W = J.shape[1:]
        dJ = du = None
        if ctx.needs_input_grad[0]:
            dJ = torch.empty((P, C, H, W), dtype=J.dtype, device=J.device)
            check(lib().b2_interp_fwd(ptr(gout), ptr(u), ptr(dJ), P, P, Pu, C, H, W, ctx.dt, ctx.bg, stream()),
                  "b2_interp_fwd")
            _lib.count_launch()
            if PJ != P:
                dJ = dJ.sum(dim=0, keepdim=True)
        if ctx.needs_input_grad[1]:
            Jx = J.expand(P, C, H, W).contiguous()
            du = torch.empty_like(u)
            check(lib().b2_interp_bwd(ptr(Jx), ptr(gout), ptr(u), None, ptr(du), P, P, Pu, C, H, W, ctx.dt, ctx.bg,
                                      stream()), "b2_interp_bwd")
            _lib.count_launch()
        return dJ, du, None, None, None


def splat(I, u, dt=1.0, need_weights=False, background="clamp"):
    """Transpose of :func:`interp` in ``I``; ``lagomorph.splat``."""
    return SplatFunction.apply(I, u, dt, BG[background], bool(need_weights))


def _field_dims(*ts):
    P, two, H, W = ts[0].shape
    for t in ts:
        if t.dim() != 4 or t.shape != ts[0].shape or t.shape[1] != 2:
            raise _lib.B2Error(f"expected equal (P,2,H,W) vector fields, got {[tuple(x.shape) for x in ts]}")
    return P, H, W


class ComposeFunction(torch.autograd.Function):
    @staticmethod
    @_lib.device_guard
    def forward(ctx, u, v, dt, background):
        u = u.contiguous()
        v = v.contiguous()
        require_cuda(u, v)
        P, H, W = _field_dims(u, v)
        out = torch.empty_like(u)
        check(lib().b2_compose_fwd(ptr(u), ptr(v), ptr(out), P, H, W, float(dt), background, stream()),
              "b2_compose_fwd")
        _lib.count_launch()
        ctx.save_for_backward(u, v)
        ctx.dt, ctx.bg = float(dt), background
        return out

    @staticmethod
    @once_differentiable
    @_lib.device_guard
    def backward(ctx, gout):
        u, v = ctx.saved_tensors
        gout = gout.contiguous()
        P, H, W = _field_dims(u, v)
        du = torch.empty_like(u) if ctx.needs_input_grad[0] else None
        dv = torch.empty_like(v) if ctx.needs_input_grad[1] else None
        check(lib().b2_compose_bwd(ptr(gout), ptr(u), ptr(v), ptr(du), ptr(dv), P, H, W, ctx.dt, ctx.bg, stream()),
              "b2_compose_bwd")
        _lib.count_launch()
        return du, dv, None, None


def compose_disp_vel(u, v, dt=1.0, background="clamp"):
    """interp(u, v, dt) + dt*v; ``lagomorph.compose_disp_vel``."""
    return ComposeFunction.apply(u, v, dt, BG[background])


class JTVFunction(torch.autograd.Function):
    @staticmethod
    @_lib.device_guard
    def forward(ctx, v, w, displacement, transpose):
        v = v.contiguous()
        w = w.contiguous()
        require_cuda(v, w)
        P, H, W = _field_dims(v, w)
        out = torch.empty_like(v)
        check(lib().b2_jtv_fwd(ptr(v), ptr(w), ptr(out), P, H, W, int(displacement), int(transpose), stream()),
              "b2_jtv_fwd")
        _lib.count_launch()
        ctx.save_for_backward(v, w)
        ctx.flags = (int(displacement), int(transpose))
        return out

    @staticmethod
    @once_differentiable
    @_lib.device_guard
    def backward(ctx, gout):
        v, w = ctx.saved_tensors
        gout = gout.contiguous()
        P, H, W = _field_dims(v, w)
        dv = torch.empty_like(v) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        check(lib().b2_jtv_bwd(ptr(gout), ptr(v), ptr(w), ptr(dv), ptr(dw), P, H, W, *ctx.flags, stream()),
              "b2_jtv_bwd")
        _lib.count_launch()
        return dv, dw, None, None


def jacobian_times_vectorfield(v, w, displacement=True, transpose=False):
    """(displacement*I + Dv) w or its transpose applied to w; ``lagomorph.jacobian_times_vectorfield``."""
    return JTVFunction.apply(v, w, displacement, transpose)


class AdStarFunction(torch.autograd.Function):
    @staticmethod
    @_lib.device_guard
    def forward(ctx, u, m, background):
        u = u.contiguous()
        m = m.contiguous()
        require_cuda(u, m)
        P, H, W = _field_dims(u, m)
        out = torch.empty_like(u)
        check(lib().b2_adstar_fwd(ptr(u), ptr(m), ptr(out), P, H, W, background, stream()), "b2_adstar_fwd")
        _lib.count_launch()
        ctx.save_for_backward(u, m)
        ctx.bg = background
        return out

    @staticmethod
    @once_differentiable
    @_lib.device_guard
    def backward(ctx, gout):
        u, m = ctx.saved_tensors
        gout = gout.contiguous()
        P, H, W = _field_dims(u, m)
        du = torch.empty_like(u) if ctx.needs_input_grad[0] else None
        dm = torch.empty_like(m) if ctx.needs_input_grad[1] else None
        ws = torch.empty_like(u) if du is not None else None
        check(lib().b2_adstar_bwd(ptr(gout), ptr(u), ptr(m), ptr(du), ptr(dm), ptr(ws), P, H, W, ctx.bg, stream()),
              "b2_adstar_bwd")
        _lib.count_launch(2 if du is not None else 1)
        return du, dm, None


def Ad_star(u, m, background="clamp"):
    """(I + Du)^T (m o (id + u)); ``lagomorph.Ad_star(u, m)``."""
    return AdStarFunction.apply(u, m, BG[background])


class FluidFunction(torch.autograd.Function):
    """flat / sharp.  Self-adjoint: the backward is the same operator on the gradient."""

    @staticmethod
    @_lib.device_guard
    def forward(ctx, f, alpha, beta, gamma, inverse):
        f = f.contiguous()
        require_cuda(f)
        if f.dim() != 4 or f.shape[1] != 2:
            raise _lib.B2Error(f"expected a (P,2,H,W) vector field, got {tuple(f.shape)}")
        ctx.params = (float(alpha), float(beta), float(gamma), int(inverse))
        return fluid_apply(f, *ctx.params)

    @staticmethod
    @once_differentiable
    @_lib.device_guard
    def backward(ctx, gout):
        return fluid_apply(gout.contiguous(), *ctx.params), None, None, None, None


def fluid_apply(f, alpha, beta, gamma, inverse):
    P, _, H, W = f.shape
    out = torch.empty_like(f)
    nbytes = lib().b2_fluid_workspace_bytes(P, H, W)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=f.device) if nbytes > 0 else None
    check(lib().b2_fluid_apply(ptr(f), ptr(out), P, H, W, alpha, beta, gamma, inverse, ptr(ws), nbytes, stream()),
          "b2_fluid_apply")
    _lib.count_launch(1 if nbytes == 0 else 3)
    return out


class FluidMetric:
    """``lagomorph.FluidMetric(params=[alpha, beta, gamma])`` with ``.flat(v)`` / ``.sharp(m)``.

    L = gamma*I - alpha*Laplacian - beta*grad(div) on the periodic grid.
    """

    def __init__(self, params=(1.0, 0.1, 0.05)):
        params = tuple(float(p) for p in params)
        if len(params) != 3:
            raise _lib.B2Error("FluidMetric expects params = [alpha, beta, gamma]")
        self.alpha, self.beta, self.gamma = params
        if not self.gamma > 0:
            raise _lib.B2Error("FluidMetric needs gamma > 0")
        self.params = params

    def flat(self, v):
        return FluidFunction.apply(v, self.alpha, self.beta, self.gamma, 0)

    def sharp(self, m):
        return FluidFunction.apply(m, self.alpha, self.beta, self.gamma, 1)

    def __repr__(self):
        return f"FluidMetric(alpha={self.alpha}, beta={self.beta}, gamma={self.gamma})"

"""Torch-CPU restatement of the lagomorph-style LDDMM operators (TEST INFRASTRUCTURE).

Every function follows SURVEY.md Appendix A (the reference tree only *imports*
lagomorph, e.g. /root/reference/modules/trainer/joint_registration_strainmat_LMA.py:5,
and never calls it in-tree; the call sites that consume the results are cited
per function).  All ops are written with differentiable torch primitives so
``torch.autograd`` through the oracle is the reference for every adjoint, and
they are dtype-generic (float64 for ground truth / gradcheck, float32 for the
parity comparisons and for the CPU baseline).

Layout: fields are (P, C, H, W); vector fields have C = 2 with component 0
along rows (H) and component 1 along columns (W); displacements are in pixels.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


@dataclass(frozen=True)
class Conventions:
    """The frozen open decisions of SURVEY.md section 8c (D1-D5)."""

    background: str = "clamp"      # D1: 'clamp' (default) | 'zero'
    edge_diff: str = "onesided"    # D2: 'onesided' (default) | 'periodic'
    adstar_det: bool = False       # D3: multiply Ad* by det(I + Du)


DEFAULT = Conventions()


# --------------------------------------------------------------------------- #
# A.1 / A.2  bilinear interp and its transpose
# --------------------------------------------------------------------------- #
def _taps(u: torch.Tensor, dt: float, background: str):
    """Tap indices and weights for sampling at x + dt*u(x).

    Returns four (flat_index (Pu,H*W) int64, weight (Pu,H*W)) pairs.
    """
    Pu, two, H, W = u.shape
    assert two == 2
    rr = torch.arange(H, dtype=u.dtype, device=u.device).view(1, H, 1)
    cc = torch.arange(W, dtype=u.dtype, device=u.device).view(1, 1, W)
    p0 = rr + dt * u[:, 0]
    p1 = cc + dt * u[:, 1]
    f0 = torch.floor(p0)
    f1 = torch.floor(p1)
    a = p0 - f0
    b = p1 - f1
    # keep the integer conversion safe for wild displacements
    i0 = f0.detach().clamp(-2.0, H + 1.0).to(torch.int64)
    j0 = f1.detach().clamp(-2.0, W + 1.0).to(torch.int64)
    taps = []
    for di, wi in ((0, 1.0 - a), (1, a)):
        for dj, wj in ((0, 1.0 - b), (1, b)):
            ii = i0 + di
            jj = j0 + dj
            w = wi * wj
            if background == "zero":
                ok = (ii >= 0) & (ii < H) & (jj >= 0) & (jj < W)
                w = w * ok.to(w.dtype)
            elif background != "clamp":
                raise ValueError(f"unknown background rule {background!r}")
            ii = ii.clamp(0, H - 1)
            jj = jj.clamp(0, W - 1)
            taps.append(((ii * W + jj).reshape(Pu, H * W), w.reshape(Pu, H * W)))
    return taps


def _bcast(I: torch.Tensor, u: torch.Tensor):
    PI, Pu = I.shape[0], u.shape[0]
    P = max(PI, Pu)
    if PI not in (1, P) or Pu not in (1, P):
        raise ValueError(f"batch sizes {PI} and {Pu} do not broadcast")
    return P


def interp(I: torch.Tensor, u: torch.Tensor, dt: float = 1.0, conv: Conventions = DEFAULT):
    """out(x) = I(x + dt*u(x)), bilinear (lagomorph.interp).

    I: (P|1, C, H, W), u: (P|1, 2, H, W) -> (P, C, H, W).  The output feeds
    'deformed_source' (joint_registration_strainmat_LMA.py:315).
    """
    P = _bcast(I, u)
    C, H, W = I.shape[1:]
    assert u.shape[-2:] == (H, W)
    Iflat = I.reshape(I.shape[0], C, H * W).expand(P, C, H * W)
    out = None
    for idx, w in _taps(u, dt, conv.background):
        idx = idx.expand(P, H * W).unsqueeze(1).expand(P, C, H * W)
        term = torch.gather(Iflat, 2, idx) * w.expand(P, H * W).unsqueeze(1)
        out = term if out is None else out + term
    return out.reshape(P, C, H, W)


def splat(J: torch.Tensor, u: torch.Tensor, dt: float = 1.0, need_weights: bool = False,
          conv: Conventions = DEFAULT):
    """Transpose of :func:`interp` in ``I`` (lagomorph.splat)."""
    P = _bcast(J, u)
    C, H, W = J.shape[1:]
    Jflat = J.reshape(J.shape[0], C, H * W).expand(P, C, H * W)
    out = torch.zeros(P, C, H * W, dtype=J.dtype, device=J.device)
    wout = torch.zeros(P, 1, H * W, dtype=J.dtype, device=J.device) if need_weights else None
    for idx, w in _taps(u, dt, conv.background):
        idx = idx.expand(P, H * W)
        w = w.expand(P, H * W)
        out = out.scatter_add(2, idx.unsqueeze(1).expand(P, C, H * W), Jflat * w.unsqueeze(1))
        if need_weights:
            wout = wout.scatter_add(2, idx.unsqueeze(1), w.unsqueeze(1))
    out = out.reshape(P, C, H, W)
    if need_weights:
        return out, wout.reshape(P, 1, H, W)
    return out


# --------------------------------------------------------------------------- #
# A.3  finite differences / Jacobian
# --------------------------------------------------------------------------- #
def _diff(f: torch.Tensor, dim: int, mode: str):
    """Central difference along ``dim``; one-sided or periodic at the ends (A.3)."""
    n = f.shape[dim]
    if mode == "periodic":
        return 0.5 * (torch.roll(f, -1, dim) - torch.roll(f, 1, dim))
    if mode != "onesided":
        raise ValueError(f"unknown edge rule {mode!r}")
    if n == 1:
        return torch.zeros_like(f)
    first = f.narrow(dim, 1, 1) - f.narrow(dim, 0, 1)
    last = f.narrow(dim, n - 1, 1) - f.narrow(dim, n - 2, 1)
    if n == 2:
        return torch.cat([first, last], dim)
    mid = 0.5 * (f.narrow(dim, 2, n - 2) - f.narrow(dim, 0, n - 2))
    return torch.cat([first, mid, last], dim)


def jacobian(v: torch.Tensor, conv: Conventions = DEFAULT):
    """(Dv)_{ab} = d v_a / d x_b as a (P, 2, 2, H, W) tensor."""
    d0 = _diff(v, 2, conv.edge_diff)   # d/d row
    d1 = _diff(v, 3, conv.edge_diff)   # d/d col
    return torch.stack([d0, d1], dim=2)


def jacobian_times_vectorfield(v, w, displacement: bool = True, transpose: bool = False,
                               conv: Conventions = DEFAULT):
    """(delta*I + Dv) w  or its transpose applied to w (lagomorph op)."""
    D = jacobian(v, conv)
    if transpose:
        D = D.transpose(1, 2)
    out = (D * w.unsqueeze(1)).sum(dim=2)
    if displacement:
        out = out + w
    return out


def Ad_star(u, m, conv: Conventions = DEFAULT):
    """Ad*_{phi^-1} m = (I + Du)^T (m o (id + u))  (SURVEY.md A.4, D3)."""
    w = interp(m, u, 1.0, conv)
    out = jacobian_times_vectorfield(u, w, displacement=True, transpose=True, conv=conv)
    if conv.adstar_det:
        D = jacobian(u, conv)
        det = (1.0 + D[:, 0, 0]) * (1.0 + D[:, 1, 1]) - D[:, 0, 1] * D[:, 1, 0]
        out = out * det.unsqueeze(1)
    return out


def compose_disp_vel(u, v, dt: float = 1.0, conv: Conventions = DEFAULT):
    """interp(u, v, dt) + dt*v  (lagomorph.compose_disp_vel)."""
    return interp(u, v, dt, conv) + dt * v


# --------------------------------------------------------------------------- #
# A.5  fluid metric
# --------------------------------------------------------------------------- #
class FluidMetric:
    """L = gamma*I - alpha*Lap_h - beta*grad_h(div_h) on a periodic grid.

    ``flat`` multiplies the orthonormal rfft2 spectrum by the per-frequency SPD
    2x2 symbol, ``sharp`` solves with it.  Both are self-adjoint, so autograd
    through them is the operator itself.
    """

    def __init__(self, params=(1.0, 0.1, 0.05)):
        self.alpha, self.beta, self.gamma = (float(p) for p in params)
        if not self.gamma > 0:
            raise ValueError("gamma must be > 0 (DC bin)")
        self._lut = {}

    def luts(self, H, W, dtype, device):
        """cos/sin lookup tables, computed in float64 then cast (upstream uses LUTs too)."""
        key = (H, W, dtype, str(device))
        if key not in self._lut:
            k0 = torch.arange(H, dtype=torch.float64)
            k1 = torch.arange(W // 2 + 1, dtype=torch.float64)
            c0 = (2.0 * (1.0 - torch.cos(2.0 * math.pi * k0 / H))).to(dtype).to(device)
            s0 = torch.sin(2.0 * math.pi * k0 / H).to(dtype).to(device)
            c1 = (2.0 * (1.0 - torch.cos(2.0 * math.pi * k1 / W))).to(dtype).to(device)
            s1 = torch.sin(2.0 * math.pi * k1 / W).to(dtype).to(device)
            self._lut[key] = (c0.view(H, 1), s0.view(H, 1), c1.view(1, -1), s1.view(1, -1))
        return self._lut[key]

    def symbol(self, H, W, dtype, device):
        c0, s0, c1, s1 = self.luts(H, W, dtype, device)
        lam = self.gamma + self.alpha * (c0 + c1)
        L00 = lam + self.beta * c0
        L11 = lam + self.beta * c1
        L01 = self.beta * (s0 * s1)
        return L00, L01, L11

    def _apply(self, f, inverse):
        P, two, H, W = f.shape
        assert two == 2
        L00, L01, L11 = self.symbol(H, W, f.dtype, f.device)
        F = torch.fft.rfft2(f, norm="ortho")
        F0, F1 = F[:, 0], F[:, 1]
        if inverse:
            det = L00 * L11 - L01 * L01
            G0 = (L11 * F0 - L01 * F1) / det
            G1 = (L00 * F1 - L01 * F0) / det
        else:
            G0 = L00 * F0 + L01 * F1
            G1 = L01 * F0 + L11 * F1
        return torch.fft.irfft2(torch.stack([G0, G1], dim=1), s=(H, W), norm="ortho")

    def flat(self, v):
        return self._apply(v, inverse=False)

    def sharp(self, m):
        return self._apply(m, inverse=True)


# --------------------------------------------------------------------------- #
# A.6  geodesic shooting
# --------------------------------------------------------------------------- #
def EPDiff_step(metric: FluidMetric, m0, dt: float, phiinv, conv: Conventions = DEFAULT,
                return_mv: bool = False, mommask=None):
    """One step of lagomorph.EPDiff_step; the optional momentum mask multiplies the
    transported momentum before ``sharp``."""
    m = Ad_star(phiinv, m0, conv)
    if mommask is not None:
        m = m * mommask
    v = metric.sharp(m)
    new = compose_disp_vel(phiinv, v, -dt, conv)
    if return_mv:
        return new, m, v
    return new


def expmap(metric: FluidMetric, m0, T: float = 1.0, num_steps: int = 10, phiinv=None,
           conv: Conventions = DEFAULT, trajectory: bool = False, mommask=None):
    """lagomorph.expmap: inverse-map displacement u with phi^-1(x) = x + u(x) (D4)."""
    u = torch.zeros_like(m0) if phiinv is None else phiinv
    dt = T / num_steps
    traj = []
    for _ in range(num_steps):
        if trajectory:
            u_prev = u
            u, m, v = EPDiff_step(metric, m0, dt, u_prev, conv, return_mv=True, mommask=mommask)
            traj.append((u_prev, m, v))
        else:
            u = EPDiff_step(metric, m0, dt, u, conv, mommask=mommask)
    if trajectory:
        return u, traj
    return u


def shoot(metric: FluidMetric, v0, num_steps: int = 10, T: float = 1.0, conv: Conventions = DEFAULT):
    """v0 -> (m0 = flat(v0), velocity = sharp(m0), displacement u^S)  (A.6, D5)."""
    m0 = metric.flat(v0)
    vel = metric.sharp(m0)
    u = expmap(metric, m0, T, num_steps, conv=conv)
    return m0, vel, u

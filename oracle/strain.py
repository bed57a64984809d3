"""Torch/numpy CPU restatement of the strain + 126-sector reduction (TEST INFRASTRUCTURE).

[SPEC] rows 17-18 of SURVEY.md section 8a: no implementation exists in the
reference tree.  The only in-tree definition of the 126 sectors is the polar
mesh of /root/reference/modules/data/utils/DENSE_utils.py:177-295 (Nseg=18 x
Nperseg=7 = 126 angular samples) together with the rotation equivariance of
/root/reference/modules/data/augmentation/affine.py:52-87 (rotating the image
by -n*360/126 degrees rolls the strain rows by +n), which fixes the index
direction: sector = floor(theta / (2*pi/126)) with theta = atan2(d_row, d_col).

Sector frame of a slice: the reference's mesh starts at theta0 = arctan2(PositionB - PositionA)
(DENSE_utils.py:198) and numbers its sectors clockwise (theta growing) or counter-clockwise per subject
(``Clockwise`` flag, DENSE_utils.py:201-204).  ``theta0`` rotates the boundary table; ``clockwise=False`` maps
index k -> n-1-k.  ``tests/golden/ref_sectors.npz`` (made by importing the reference's own ``spl2patchSA``) pins
start angle, direction and count.

Sector assignment is integer-only (D7) so CPU and GPU agree bit for bit:
with cnt = sum(mask0), sx = sum(row*mask0), sy = sum(col*mask0) the direction of
pixel (r, c) from the centroid is d = (cnt*r - sx, cnt*c - sy) (int64), and
sector k is the unique wedge with cross(b_k, d) >= 0 > cross(b_{k+1}, d), where
b_k = (round(2^20 sin(2 pi k/n)), round(2^20 cos(2 pi k/n))).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .lddmm import Conventions, DEFAULT, jacobian

N_SECTORS = 126
Q = 1 << 20
DET_EPS = 1e-6
RAD2_EPS = 1e-12


def sector_boundaries(n_sectors: int = N_SECTORS, theta0: float = 0.0) -> np.ndarray:
    """(n, 2) int64 table of Q20 boundary directions (row, col), rotated by ``theta0`` radians."""
    k = np.arange(n_sectors, dtype=np.float64)
    ang = float(theta0) + 2.0 * math.pi * k / n_sectors
    br = np.rint(Q * np.sin(ang)).astype(np.int64)
    bc = np.rint(Q * np.cos(ang)).astype(np.int64)
    return np.stack([br, bc], axis=1)


def mask_moments(mask0: torch.Tensor):
    """cnt, sx, sy (int64, per slice) of the frame-0 mask ``mask0 > 0.5``  (B, H, W)."""
    m = (mask0 > 0.5).to(torch.int64)
    B, H, W = m.shape
    rr = torch.arange(H, dtype=torch.int64).view(1, H, 1)
    cc = torch.arange(W, dtype=torch.int64).view(1, 1, W)
    cnt = m.sum(dim=(1, 2))
    sx = (m * rr).sum(dim=(1, 2))
    sy = (m * cc).sum(dim=(1, 2))
    return cnt, sx, sy


def centroid(cnt, sx, sy, H, W, dtype):
    """Centroid as float: double division then cast; image centre for an empty mask."""
    cntd = cnt.to(torch.float64)
    c0 = torch.where(cnt > 0, sx.to(torch.float64) / cntd.clamp(min=1.0),
                     torch.full_like(cntd, (H - 1) / 2.0))
    c1 = torch.where(cnt > 0, sy.to(torch.float64) / cntd.clamp(min=1.0),
                     torch.full_like(cntd, (W - 1) / 2.0))
    return c0.to(dtype), c1.to(dtype)


def classify_directions(dr: np.ndarray, dc: np.ndarray, n_sectors: int = N_SECTORS, theta0: float = 0.0,
                        clockwise: bool = True) -> np.ndarray:
    """Integer-exact sector of int64 directions (dr, dc) in the frame (theta0, clockwise); -1 for the zero vector."""
    tab = sector_boundaries(n_sectors, theta0)
    br, bc = tab[:, 0], tab[:, 1]
    dr = np.asarray(dr, dtype=np.int64)
    dc = np.asarray(dc, dtype=np.int64)
    theta = np.arctan2(dr.astype(np.float64), dc.astype(np.float64)) - float(theta0)
    theta = np.mod(theta, 2.0 * math.pi)
    k = np.floor(theta / (2.0 * math.pi / n_sectors)).astype(np.int64) % n_sectors
    zero = (dr == 0) & (dc == 0)
    for _ in range(n_sectors):
        k1 = (k + 1) % n_sectors
        lo = bc[k] * dr - br[k] * dc
        hi = bc[k1] * dr - br[k1] * dc
        down = (lo < 0) & ~zero
        up = (hi >= 0) & ~down & ~zero
        if not (down.any() or up.any()):
            break
        k = np.where(down, (k - 1) % n_sectors, np.where(up, k1, k))
    else:  # pragma: no cover - cannot happen for a valid boundary table
        raise RuntimeError("sector classification did not converge")
    if not clockwise:
        k = n_sectors - 1 - k
    return np.where(zero, -1, k).astype(np.int32)


def _per_slice(x, B, default):
    if x is None:
        return [default] * B
    a = np.asarray(x).reshape(-1)
    if a.size == 1:
        a = np.repeat(a, B)
    assert a.size == B, f"{a.size} entries for {B} slices"
    return list(a)


def sector_map(mask0: torch.Tensor, n_sectors: int = N_SECTORS, theta0=None, clockwise=None) -> torch.Tensor:
    """(B, H, W) int32 sector id of every pixel about the frame-0 mask centroid; ``theta0`` (radians) and
    ``clockwise`` are scalars or one entry per slice (default 0 / True)."""
    B, H, W = mask0.shape
    cnt, sx, sy = mask_moments(mask0)
    rr = np.arange(H, dtype=np.int64).reshape(H, 1)
    cc = np.arange(W, dtype=np.int64).reshape(1, W)
    th, cw = _per_slice(theta0, B, 0.0), _per_slice(clockwise, B, True)
    out = np.empty((B, H, W), np.int32)
    for b in range(B):
        dr = int(cnt[b]) * rr - int(sx[b]) + 0 * cc
        dc = int(cnt[b]) * cc - int(sy[b]) + 0 * rr
        out[b] = classify_directions(dr, dc, n_sectors, float(th[b]), bool(cw[b]))
    return torch.from_numpy(out)


def strain_ecc(u: torch.Tensor, ctr0: torch.Tensor, ctr1: torch.Tensor, conv: Conventions = DEFAULT):
    """Circumferential Green-Lagrange strain of the inverse-map displacement (A.7).

    u: (P, 2, H, W); ctr0/ctr1: (P,) centroid of the slice's frame-0 mask.
    G = I + Du, F = G^-1, Ecc = 0.5*(|F e_theta|^2 - 1) with e_theta perpendicular to
    X - ctr, X = x + u(x).  Returns (Ecc (P,H,W), valid (P,H,W) bool).
    """
    P, _, H, W = u.shape
    D = jacobian(u, conv)
    G00 = 1.0 + D[:, 0, 0]
    G01 = D[:, 0, 1]
    G10 = D[:, 1, 0]
    G11 = 1.0 + D[:, 1, 1]
    det = G00 * G11 - G01 * G10
    rr = torch.arange(H, dtype=u.dtype).view(1, H, 1)
    cc = torch.arange(W, dtype=u.dtype).view(1, 1, W)
    n0 = (rr + u[:, 0]) - ctr0.view(P, 1, 1)
    n1 = (cc + u[:, 1]) - ctr1.view(P, 1, 1)
    rad2 = n0 * n0 + n1 * n1
    e0, e1 = -n1, n0                       # unnormalised e_theta
    t0 = G11 * e0 - G01 * e1               # adj(G) e
    t1 = G00 * e1 - G10 * e0
    valid = (rad2 >= RAD2_EPS) & (det.abs() >= DET_EPS)
    den = torch.where(valid, rad2 * det * det, torch.ones_like(det))
    ecc = 0.5 * ((t0 * t0 + t1 * t1) / den - 1.0)
    return torch.where(valid, ecc, torch.zeros_like(ecc)), valid


def align_frames(S: torch.Tensor, n_frames: int) -> torch.Tensor:
    """Crop or edge-pad the last (frame) axis to ``n_frames``.

    Same rule as /root/reference/modules/data/datareader/DENSE_IO_utils.py:26-46.
    """
    T = S.shape[-1]
    if T >= n_frames:
        return S[..., :n_frames]
    pad = S[..., -1:].expand(*S.shape[:-1], n_frames - T)
    return torch.cat([S, pad], dim=-1)


def strain_matrix(u: torch.Tensor, tar: torch.Tensor, mask0: torch.Tensor,
                  n_sectors: int = N_SECTORS, n_frames: int | None = 40,
                  conv: Conventions = DEFAULT, return_counts: bool = False, theta0=None, clockwise=None):
    """Masked per-sector mean of Ecc.

    u: (B, T1, 2, H, W) inverse-map displacements of the T1 = T-1 frame-pairs of
    each slice; tar: (B, T1, H, W) target-frame masks; mask0: (B, H, W) frame-0 mask.
    Returns (B, 1, n_sectors, n_frames) - the layout of the ground-truth strain
    matrix (/root/reference/modules/data/dataset/joint_dataset.py:72).
    """
    B, T1, _, H, W = u.shape
    cnt, sx, sy = mask_moments(mask0)
    c0, c1 = centroid(cnt, sx, sy, H, W, u.dtype)
    sect = sector_map(mask0, n_sectors, theta0, clockwise)    # (B,H,W) int32
    ecc, valid = strain_ecc(u.reshape(B * T1, 2, H, W),
                            c0.repeat_interleave(T1), c1.repeat_interleave(T1), conv)
    ecc = ecc.reshape(B, T1, H * W)
    member = (tar.reshape(B, T1, H * W) > 0.5) & valid.reshape(B, T1, H * W) \
        & (sect.reshape(B, 1, H * W) >= 0)
    idx = sect.reshape(B, 1, H * W).clamp(min=0).to(torch.int64).expand(B, T1, H * W)
    memf = member.to(u.dtype)
    sums = torch.zeros(B, T1, n_sectors, dtype=u.dtype).scatter_add(2, idx, ecc * memf)
    cnts = torch.zeros(B, T1, n_sectors, dtype=u.dtype).scatter_add(2, idx, memf)
    S = (sums / cnts.clamp(min=1.0)).permute(0, 2, 1).unsqueeze(1)   # (B,1,K,T1)
    if n_frames is not None:
        S = align_frames(S, n_frames)
    if return_counts:
        return S, cnts.permute(0, 2, 1).to(torch.int32)
    return S

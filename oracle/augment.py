"""Oracle for the augmentation step in front of the path (TEST INFRASTRUCTURE).

Restates /root/reference/modules/data/augmentation/affine.py:
* ``translate`` (:24-50)  - ``np.roll`` of the masks by (translate_y, translate_x) on the (row, col) axes; strain
  matrix and TOS unchanged.  Golden-checked against the reference's own function.
* ``rotate`` (:52-87)     - ``skimage.transform.rotate(mask, -n*360/126, resize=False, preserve_range=True, order=0)``
  and ``np.roll(strain, n, axis=0)``, ``np.roll(TOS, n)``.  The rolls and the angle/keyword arguments handed to
  skimage are golden-checked; skimage itself (third-party, absent here, unpinned) is restated from its published
  algorithm: inverse map  in = c + R(theta) (out - c)  in (col, row) coordinates, c = (W/2 - 0.5, H/2 - 0.5),
  theta = deg2rad(angle), nearest neighbour = C ``round`` (half away from zero), outside -> 0 (mode='constant').
  "Parity unpinned" for that one call.
* order of application: rotate, then translate (/root/reference/modules/data/augmentation/__init__.py:20-21).

Coordinates are float64 with separately rounded multiplies and adds (numpy semantics), which the CUDA kernel
reproduces with __dmul_rn/__dadd_rn, so the selected source index - integer work - is bit-exact.
"""
from __future__ import annotations

import math

import numpy as np


def rotation_angle_degree(n_rotate_sectors: int, n_total_sectors: int = 126) -> float:
    """affine.py:56."""
    return -n_rotate_sectors * 360 / n_total_sectors


def rotate_matrix(angle_degree: float, H: int, W: int) -> np.ndarray:
    """(2,3) float64 map from output (col, row, 1) to input (col, row): skimage.transform.rotate's
    ``tform3 + tform2 + tform1`` (translate by -c, rotate, translate by +c), multiplied out in float64."""
    th = math.radians(angle_degree)
    cs, sn = math.cos(th), math.sin(th)
    cx, cy = W / 2.0 - 0.5, H / 2.0 - 0.5
    return np.array([[cs, -sn, cx - (cs * cx - sn * cy)],
                     [sn, cs, cy - (sn * cx + cs * cy)]], dtype=np.float64)


def _round_half_away(x: np.ndarray) -> np.ndarray:
    return np.where(x >= 0, np.floor(x + 0.5), np.ceil(x - 0.5))


def rotate_nearest(img: np.ndarray, matrix: np.ndarray) -> np.ndarray:
    """img (..., H, W); nearest-neighbour warp with the inverse map ``matrix``; outside -> 0."""
    H, W = img.shape[-2:]
    r, c = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    x_in = (matrix[0, 0] * c + matrix[0, 1] * r) + matrix[0, 2]
    y_in = (matrix[1, 0] * c + matrix[1, 1] * r) + matrix[1, 2]
    ci = _round_half_away(x_in).astype(np.int64)
    ri = _round_half_away(y_in).astype(np.int64)
    ok = (ci >= 0) & (ci < W) & (ri >= 0) & (ri < H)
    out = np.zeros_like(img)
    out[..., ok] = img[..., ri[ok], ci[ok]]
    return out


def rotate_translate_volume(vol: np.ndarray, n_rotate_sectors, translate_y, translate_x,
                            n_total_sectors: int = 126) -> np.ndarray:
    """vol (B,1,T,H,W); per-slice integer arrays (B,).  rotate, then circular roll (rows by ty, cols by tx)."""
    B = vol.shape[0]
    H, W = vol.shape[-2:]
    out = np.empty_like(vol)
    for b in range(B):
        m = rotate_matrix(rotation_angle_degree(int(n_rotate_sectors[b]), n_total_sectors), H, W)
        rot = rotate_nearest(vol[b], m)
        out[b] = np.roll(rot, (int(translate_y[b]), int(translate_x[b])), axis=(-2, -1))
    return out


def roll_rows(S: np.ndarray, n) -> np.ndarray:
    """S (B, ..., R, C) or TOS (B, R): ``np.roll(x, n_b, axis=row axis)`` per sample (affine.py:74,78).
    The row axis is -2 for matrices (B,1,126,40) and -1 for (B,126) vectors."""
    out = np.empty_like(S)
    ax = -1 if S.ndim == 2 else -2
    for b in range(S.shape[0]):
        out[b] = np.roll(S[b], int(n[b]), axis=ax)
    return out

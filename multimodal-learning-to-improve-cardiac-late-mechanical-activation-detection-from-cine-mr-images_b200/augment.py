"""Device-side augmentation of a cine batch.

Mirrors /root/reference/modules/data/augmentation/affine.py: ``rotate`` by ``n`` sectors
(``skimage.transform.rotate(mask, -n*360/126, resize=False, preserve_range=True, order=0)`` plus
``np.roll(strain, n, axis=0)`` / ``np.roll(TOS, n)``, :52-87) followed by ``translate`` (``np.roll`` of the masks by
``(translate_y, translate_x)``, strain and TOS unchanged, :24-50) - the order of
augmentation/__init__.py:20-21 - on the (B,1,T,H,W) batch the path consumes, one transform per slice, in one
gather kernel (``b2_augment_volume``) and one row-roll kernel (``b2_roll_rows``).
"""
from __future__ import annotations

import math

import torch

from . import _lib
from ._lib import check, lib, ptr, require_cuda, stream


def rotation_matrix(n_rotate_sectors: int, H: int, W: int, n_total_sectors: int = 126):
    """Six float64 numbers of skimage's inverse map (output (col,row,1) -> input (col,row)) for ``rotate(datum, n)``."""
    th = math.radians(-n_rotate_sectors * 360 / n_total_sectors)       # affine.py:56
    cs, sn = math.cos(th), math.sin(th)
    cx, cy = W / 2.0 - 0.5, H / 2.0 - 0.5                              # skimage: centre = (cols, rows) / 2 - 0.5
    return [cs, -sn, cx - (cs * cx - sn * cy), sn, cs, cy - (sn * cx + cs * cy)]


def _per_slice(x, B, name):
    if x is None:
        return [0] * B
    if isinstance(x, int):
        return [x] * B
    x = [int(v) for v in (x.tolist() if hasattr(x, "tolist") else x)]
    if len(x) != B:
        raise _lib.B2Error(f"{name}: expected {B} per-slice values, got {len(x)}")
    return x


def rotate_translate_volume(vol, n_rotate_sectors=0, translate_y=0, translate_x=0, n_total_sectors=126):
    """vol (B,1,T,H,W) fp32 CUDA; per-slice ints (or one int for all).  Returns the augmented volume."""
    vol = vol.contiguous()
    require_cuda(vol)
    if vol.dim() != 5 or vol.shape[1] != 1:
        raise _lib.B2Error(f"expected a (B,1,T,H,W) volume, got {tuple(vol.shape)}")
    B, _, T, H, W = vol.shape
    ns = _per_slice(n_rotate_sectors, B, "n_rotate_sectors")
    ty = _per_slice(translate_y, B, "translate_y")
    tx = _per_slice(translate_x, B, "translate_x")
    xf = torch.tensor([rotation_matrix(n, H, W, n_total_sectors) for n in ns], dtype=torch.float64).to(vol.device)
    sh = torch.tensor(list(zip(ty, tx)), dtype=torch.int32).to(vol.device)
    out = torch.empty_like(vol)
    check(lib().b2_augment_volume(ptr(vol), ptr(out), ptr(xf), ptr(sh), B, T, H, W, stream()), "b2_augment_volume")
    _lib.count_launch()
    return out


def roll_rows(S, n):
    """Strain matrices (B,1,R,C) / (B,R,C) or TOS curves (B,R): rows rolled by ``n[b]`` (np.roll(x, n, axis=0))."""
    S = S.contiguous()
    require_cuda(S)
    B = S.shape[0]
    if S.dim() == 2:
        R, Cc = S.shape[1], 1
    elif S.dim() in (3, 4) and (S.dim() == 3 or S.shape[1] == 1):
        R, Cc = S.shape[-2], S.shape[-1]
    else:
        raise _lib.B2Error(f"expected (B,R), (B,R,C) or (B,1,R,C), got {tuple(S.shape)}")
    nn = torch.tensor(_per_slice(n, B, "n"), dtype=torch.int32).to(S.device)
    out = torch.empty_like(S)
    check(lib().b2_roll_rows(ptr(S), ptr(out), ptr(nn), B, R, Cc, stream()), "b2_roll_rows")
    _lib.count_launch()
    return out


def augment_batch(vol, strain_matrix=None, TOS=None, n_rotate_sectors=0, translate_y=0, translate_x=0,
                  n_total_sectors=126):
    """``augment_datum`` of the reference (augmentation/__init__.py:4-23) for a whole device batch."""
    out = {"cine_myo_mask": rotate_translate_volume(vol, n_rotate_sectors, translate_y, translate_x, n_total_sectors)}
    if strain_matrix is not None:
        out["strain_matrix"] = roll_rows(strain_matrix, n_rotate_sectors)
    if TOS is not None:
        out["TOS"] = roll_rows(TOS, n_rotate_sectors)
    return out

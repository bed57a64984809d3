"""Strain matrix: per-slice deformation-gradient strain + 126-sector masked mean.

[SPEC] rows 17-18 of SURVEY.md section 8a.  Output layout (B, 1, n_sectors,
n_frames) matches the ground-truth strain matrix the trainer compares with
(/root/reference/modules/data/dataset/joint_dataset.py:72 under
``MSELoss``, /root/reference/modules/loss/loss_calculator.py:65-67).
"""
from __future__ import annotations

import ctypes as C

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import check, lib, ptr, require_cuda, stream

N_SECTORS = 126
_table_cache = {}


def sector_table(n_sectors: int, device) -> torch.Tensor:
    """(n_sectors, 2) int32 Q20 boundary directions on ``device`` (host-computed, cached)."""
    key = (int(n_sectors), str(device))
    if key not in _table_cache:
        buf = (C.c_int32 * (2 * n_sectors))()
        check(lib().b2_sector_table_host(int(n_sectors), buf), "b2_sector_table_host")
        host = torch.tensor(list(buf), dtype=torch.int32).view(n_sectors, 2)
        _table_cache[key] = host.to(device)
    return _table_cache[key]


def mask_moments(mask0: torch.Tensor) -> torch.Tensor:
    """(B,3) int64 {count, sum(row), sum(col)} of ``mask0 > 0.5``; mask0 is (B,H,W)."""
    mask0 = mask0.contiguous()
    require_cuda(mask0)
    B, H, W = mask0.shape
    mom = torch.empty((B, 3), dtype=torch.int64, device=mask0.device)
    check(lib().b2_mask_moments(ptr(mask0), ptr(mom), B, H, W, stream()), "b2_mask_moments")
    _lib.count_launch()
    return mom


def sector_map(mask0: torch.Tensor, n_sectors: int = N_SECTORS) -> torch.Tensor:
    """(B,H,W) int32 sector id of every pixel about the frame-0 mask centroid (-1 at the centroid)."""
    mom = mask_moments(mask0)
    B, H, W = mask0.shape
    out = torch.empty((B, H, W), dtype=torch.int32, device=mask0.device)
    check(lib().b2_sector_map_i32(ptr(mom), ptr(sector_table(n_sectors, mask0.device)), ptr(out), B, H, W,
                                  n_sectors, stream()), "b2_sector_map_i32")
    _lib.count_launch()
    return out


class StrainMatrixFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, tar, moments, table, n_sectors, n_frames):
        u = u.contiguous()
        tar = tar.contiguous()
        require_cuda(u, tar)
        B, T1, two, H, W = u.shape
        if two != 2 or tuple(tar.shape) != (B, T1, H, W):
            raise _lib.B2Error(f"expected u (B,T1,2,H,W) and tar (B,T1,H,W), got {tuple(u.shape)}, {tuple(tar.shape)}")
        S = torch.empty((B, 1, n_sectors, n_frames), dtype=u.dtype, device=u.device)
        counts = torch.empty((B, n_sectors, T1), dtype=torch.int32, device=u.device)
        check(lib().b2_strain_sector_fwd(ptr(u), ptr(tar), ptr(moments), ptr(table), ptr(S), ptr(counts), B, T1, H, W,
                                         n_sectors, n_frames, stream()), "b2_strain_sector_fwd")
        _lib.count_launch()
        ctx.save_for_backward(u, tar, moments, table, counts)
        ctx.dims = (n_sectors, n_frames)
        ctx.mark_non_differentiable(counts)
        return S, counts

    @staticmethod
    @once_differentiable
    def backward(ctx, gS, _gc=None):
        u, tar, moments, table, counts = ctx.saved_tensors
        B, T1, _, H, W = u.shape
        du = torch.empty_like(u)
        gS_c = gS.contiguous()
        check(lib().b2_strain_sector_bwd(ptr(gS_c), ptr(u), ptr(tar), ptr(moments), ptr(table),
                                         ptr(counts), ptr(du), B, T1, H, W, *ctx.dims, stream()),
              "b2_strain_sector_bwd")
        _lib.count_launch()
        return du, None, None, None, None, None


def strain_matrix(u, tar, mask0, n_sectors: int = N_SECTORS, n_frames: int | None = 40, return_counts=False):
    """Masked per-sector mean circumferential strain.

    u: (B,T1,2,H,W) inverse-map displacements; tar: (B,T1,H,W) target masks;
    mask0: (B,H,W) frame-0 mask.  Returns (B,1,n_sectors,n_frames).
    """
    if n_frames is None:
        n_frames = u.shape[1]
    mom = mask_moments(mask0)
    S, counts = StrainMatrixFunction.apply(u, tar, mom, sector_table(n_sectors, u.device), int(n_sectors),
                                           int(n_frames))
    return (S, counts) if return_counts else S

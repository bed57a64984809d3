"""Strain matrix: per-slice deformation-gradient strain + 126-sector masked mean.

[SPEC] rows 17-18 of SURVEY.md section 8a.  Output layout (B, 1, n_sectors,
n_frames) matches the ground-truth strain matrix the trainer compares with
(/root/reference/modules/data/dataset/joint_dataset.py:72 under
``MSELoss``, /root/reference/modules/loss/loss_calculator.py:65-67).

Sector frame of a slice.  The reference's 126-sector mesh (``spl2patchSA``,
/root/reference/modules/data/utils/DENSE_utils.py:177-295) starts at
``theta0 = arctan2(PositionB - PositionA)`` (``:198``) and numbers the sectors clockwise or counter-clockwise per
subject (``Clockwise`` flag, ``:201-204``).  ``theta0`` (radians, per slice) and ``clockwise`` (bool, per slice)
carry that frame through every entry point here; the defaults (0, True) are the convention of
``augmentation/affine.py:52-87``: angle 0 = +column axis, index growing from +col toward +row.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import SectorFrame, check, lib, ptr, require_cuda, stream

N_SECTORS = 126
_table_cache = {}


def _host_table(n_sectors: int, theta0: float) -> torch.Tensor:
    buf = (C.c_int32 * (2 * n_sectors))()
    check(lib().b2_sector_table_rotated_host(int(n_sectors), float(theta0), buf), "b2_sector_table_rotated_host")
    return torch.tensor(list(buf), dtype=torch.int32).view(n_sectors, 2)


def sector_table(n_sectors: int, device) -> torch.Tensor:
    """(n_sectors, 2) int32 Q20 boundary directions on ``device`` (host-computed, cached)."""
    key = (int(n_sectors), str(device))
    if key not in _table_cache:
        _table_cache[key] = _host_table(n_sectors, 0.0).to(device)
    return _table_cache[key]


class Frame:
    """Device-side sector frame of a batch of B slices: boundary tables rotated by ``theta0`` on the host
    (integer-exact classification), the float seeds and the direction flags.  ``Frame.default`` shares one
    unrotated table between all slices."""

    def __init__(self, n_sectors: int, B: int, device, theta0=None, clockwise=None):
        self.n_sectors, self.B = int(n_sectors), int(B)
        if theta0 is None:
            self.table, self.stride, self.theta0 = sector_table(n_sectors, device), 0, None
        else:
            th = torch.as_tensor(theta0, dtype=torch.float64).reshape(-1).cpu()
            if th.numel() == 1:
                th = th.expand(B)
            if th.numel() != B:
                raise _lib.B2Error(f"theta0 has {th.numel()} entries for {B} slices")
            self.table = torch.stack([_host_table(n_sectors, float(t)) for t in th]).to(device)      # (B, n, 2)
            self.stride = 2 * self.n_sectors
            self.theta0 = th.to(torch.float32).to(device)
        if clockwise is None:
            self.clockwise = None
        else:
            cw = torch.as_tensor(clockwise).reshape(-1).cpu()
            if cw.numel() == 1:
                cw = cw.expand(B)
            if cw.numel() != B:
                raise _lib.B2Error(f"clockwise has {cw.numel()} entries for {B} slices")
            self.clockwise = (cw != 0).to(torch.int32).to(device)

    def c_struct(self, first_slice: int = 0) -> SectorFrame:
        """``b2_sector_frame`` of the slices ``first_slice ...`` (chunked launches)."""
        f = SectorFrame()
        f.table = self.table.data_ptr() + 4 * self.stride * first_slice
        f.table_slice_stride = self.stride
        f.theta0 = self.theta0.data_ptr() + 4 * first_slice if self.theta0 is not None else None
        f.clockwise = self.clockwise.data_ptr() + 4 * first_slice if self.clockwise is not None else None
        return f

    def tensors(self):
        return [t for t in (self.table, self.theta0, self.clockwise) if t is not None]


def mask_moments(mask0: torch.Tensor) -> torch.Tensor:
    """(B,3) int64 {count, sum(row), sum(col)} of ``mask0 > 0.5``; mask0 is (B,H,W)."""
    mask0 = mask0.contiguous()
    require_cuda(mask0)
    B, H, W = mask0.shape
    with _lib.on_device(mask0):
        mom = torch.empty((B, 3), dtype=torch.int64, device=mask0.device)
        check(lib().b2_mask_moments(ptr(mask0), ptr(mom), B, H, W, stream()), "b2_mask_moments")
    _lib.count_launch()
    return mom


def sector_map(mask0: torch.Tensor, n_sectors: int = N_SECTORS, theta0=None, clockwise=None) -> torch.Tensor:
    """(B,H,W) int32 sector id of every pixel about the frame-0 mask centroid (-1 at the centroid)."""
    mom = mask_moments(mask0)
    B, H, W = mask0.shape
    frame = Frame(n_sectors, B, mask0.device, theta0, clockwise)
    with _lib.on_device(mask0):
        out = torch.empty((B, H, W), dtype=torch.int32, device=mask0.device)
        fs = frame.c_struct()
        check(lib().b2_sector_map_i32_ex(ptr(mom), C.byref(fs), ptr(out), B, H, W, n_sectors, stream()),
              "b2_sector_map_i32_ex")
    _lib.count_launch()
    return out


class StrainMatrixFunction(torch.autograd.Function):
    @staticmethod
    @_lib.device_guard
    def forward(ctx, u, tar, moments, frame, n_sectors, n_frames):
        u = u.contiguous()
        tar = tar.contiguous()
        require_cuda(u, tar)
        B, T1, two, H, W = u.shape
        if two != 2 or tuple(tar.shape) != (B, T1, H, W):
            raise _lib.B2Error(f"expected u (B,T1,2,H,W) and tar (B,T1,H,W), got {tuple(u.shape)}, {tuple(tar.shape)}")
        S = torch.empty((B, 1, n_sectors, n_frames), dtype=u.dtype, device=u.device)
        counts = torch.empty((B, n_sectors, T1), dtype=torch.int32, device=u.device)
        fs = frame.c_struct()
        check(lib().b2_strain_sector_fwd_ex(ptr(u), ptr(tar), ptr(moments), C.byref(fs), ptr(S), ptr(counts), B, T1,
                                            H, W, n_sectors, n_frames, stream()), "b2_strain_sector_fwd_ex")
        _lib.count_launch()
        ctx.save_for_backward(u, tar, moments, counts)
        ctx.frame = frame
        ctx.dims = (n_sectors, n_frames)
        ctx.mark_non_differentiable(counts)
        return S, counts

    @staticmethod
    @once_differentiable
    @_lib.device_guard
    def backward(ctx, gS, _gc=None):
        u, tar, moments, counts = ctx.saved_tensors
        B, T1, _, H, W = u.shape
        du = torch.empty_like(u)
        gS_c = gS.contiguous()
        fs = ctx.frame.c_struct()
        check(lib().b2_strain_sector_bwd_ex(ptr(gS_c), ptr(u), ptr(tar), ptr(moments), C.byref(fs),
                                            ptr(counts), ptr(du), B, T1, H, W, *ctx.dims, stream()),
              "b2_strain_sector_bwd_ex")
        _lib.count_launch()
        return du, None, None, None, None, None


def strain_matrix(u, tar, mask0, n_sectors: int = N_SECTORS, n_frames: int | None = 40, return_counts=False,
                  theta0=None, clockwise=None):
    """Masked per-sector mean circumferential strain.

    u: (B,T1,2,H,W) inverse-map displacements; tar: (B,T1,H,W) target masks;
    mask0: (B,H,W) frame-0 mask.  ``theta0`` (radians) / ``clockwise``: per-slice sector frame (scalars or B
    entries; module docstring).  Returns (B,1,n_sectors,n_frames).
    """
    if n_frames is None:
        n_frames = u.shape[1]
    mom = mask_moments(mask0)
    frame = Frame(n_sectors, u.shape[0], u.device, theta0, clockwise)
    S, counts = StrainMatrixFunction.apply(u, tar, mom, frame, int(n_sectors), int(n_frames))
    return (S, counts) if return_counts else S

"""ctypes binding of the C-ABI library ``libb2lddmm.so`` (include/b2lddmm.h).

The library is the product path.  There is NO CPU fallback: if the shared
library is missing, or an op is handed a non-CUDA tensor, the call raises.
"""
from __future__ import annotations

import ctypes as C
import pathlib

import torch

import os

# B2LDDMM_LIB selects an alternative build of the same C ABI (A/B tuning runs only)
_LIB_PATH = pathlib.Path(os.environ.get("B2LDDMM_LIB") or pathlib.Path(__file__).resolve().parent / "libb2lddmm.so")
_lib = None

c_f = C.c_void_p       # float* / any device pointer
c_i64 = C.c_int64
c_int = C.c_int
c_float = C.c_float


class ShootArgs(C.Structure):
    """Mirror of ``b2_shoot_args`` (include/b2lddmm.h)."""
    _fields_ = [
        ("v0", C.c_void_p), ("src", C.c_void_p), ("tar", C.c_void_p),
        ("moments", C.c_void_p), ("table", C.c_void_p),
        ("m0", C.c_void_p), ("vel", C.c_void_p), ("u", C.c_void_p), ("sdef", C.c_void_p),
        ("S", C.c_void_p), ("counts", C.c_void_p), ("traj", C.c_void_p),
        ("B", C.c_int64), ("T1", C.c_int64), ("H", C.c_int64), ("W", C.c_int64),
        ("src_slice_stride", C.c_int64), ("tar_slice_stride", C.c_int64),
        ("num_steps", C.c_int32), ("src_per_pair", C.c_int32), ("v0_is_momentum", C.c_int32),
        ("n_sectors", C.c_int32), ("n_frames", C.c_int32), ("background", C.c_int32),
        ("alpha", C.c_float), ("beta", C.c_float), ("gamma", C.c_float), ("T", C.c_float),
        ("loss_terms", C.c_void_p),
        ("table_slice_stride", C.c_int64), ("theta0", C.c_void_p), ("clockwise", C.c_void_p),
        ("flags", C.c_int32), ("reserved_", C.c_int32),
        ("pair_begin", C.c_int64), ("pair_count", C.c_int64),
    ]


class SectorFrame(C.Structure):
    """Mirror of ``b2_sector_frame``: per-slice rotated boundary tables + search seed + direction."""
    _fields_ = [("table", C.c_void_p), ("table_slice_stride", C.c_int64), ("theta0", C.c_void_p),
                ("clockwise", C.c_void_p)]


class ShootBwdArgs(C.Structure):
    """Mirror of ``b2_shoot_bwd_args``."""
    _fields_ = [
        ("gu", C.c_void_p), ("gvel", C.c_void_p), ("gm0", C.c_void_p), ("g_reg", C.c_void_p),
        ("m0", C.c_void_p), ("traj", C.c_void_p), ("gv0", C.c_void_p),
        ("P", C.c_int64), ("H", C.c_int64), ("W", C.c_int64),
        ("num_steps", C.c_int32), ("background", C.c_int32), ("v0_is_momentum", C.c_int32), ("flags", C.c_int32),
        ("alpha", C.c_float), ("beta", C.c_float), ("gamma", C.c_float), ("T", C.c_float),
        ("seed_gS", C.c_void_p), ("seed_counts", C.c_void_p), ("seed_moments", C.c_void_p), ("seed_table", C.c_void_p),
        ("seed_table_slice_stride", C.c_int64), ("seed_theta0", C.c_void_p), ("seed_clockwise", C.c_void_p),
        ("seed_g_sq", C.c_void_p), ("seed_u", C.c_void_p), ("seed_src", C.c_void_p), ("seed_tar", C.c_void_p),
        ("seed_T1", C.c_int64), ("seed_src_slice_stride", C.c_int64), ("seed_tar_slice_stride", C.c_int64),
        ("seed_n_sectors", C.c_int32), ("seed_n_frames", C.c_int32), ("seed_src_per_pair", C.c_int32),
        ("seed_reserved_", C.c_int32),
    ]


FLAG_OPLEVEL = 1      # B2_FLAG_OPLEVEL


# name -> (restype, argtypes); kept in one table so tests can check every symbol of the header
SIGNATURES = {
    "b2_version": (c_int, []),
    "b2_error_string": (C.c_char_p, [c_int]),
    "b2_interp_fwd": (c_int, [c_f, c_f, c_f, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_float, c_int, c_f]),
    "b2_interp_bwd": (c_int, [c_f, c_f, c_f, c_f, c_f, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_float, c_int, c_f]),
    "b2_warp_fwd": (c_int, [c_f, c_f, c_f, c_i64, c_i64, c_i64, c_i64, c_i64, c_float, c_int, c_f]),
    "b2_warp_bwd": (c_int, [c_f, c_f, c_f, c_f, c_f, c_i64, c_i64, c_i64, c_i64, c_i64, c_float, c_int, c_f]),
    "b2_splat_fwd": (c_int, [c_f, c_f, c_f, c_f, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_float, c_int, c_f]),
    "b2_compose_fwd": (c_int, [c_f, c_f, c_f, c_i64, c_i64, c_i64, c_float, c_int, c_f]),
    "b2_compose_bwd": (c_int, [c_f, c_f, c_f, c_f, c_f, c_i64, c_i64, c_i64, c_float, c_int, c_f]),
    "b2_jtv_fwd": (c_int, [c_f, c_f, c_f, c_i64, c_i64, c_i64, c_int, c_int, c_f]),
    "b2_jtv_bwd": (c_int, [c_f, c_f, c_f, c_f, c_f, c_i64, c_i64, c_i64, c_int, c_int, c_f]),
    "b2_adstar_fwd": (c_int, [c_f, c_f, c_f, c_i64, c_i64, c_i64, c_int, c_f]),
    "b2_adstar_bwd": (c_int, [c_f, c_f, c_f, c_f, c_f, c_f, c_i64, c_i64, c_i64, c_int, c_f]),
    "b2_fluid_workspace_bytes": (c_i64, [c_i64, c_i64, c_i64]),
    "b2_fluid_apply": (c_int, [c_f, c_f, c_i64, c_i64, c_i64, c_float, c_float, c_float, c_int, c_f, c_i64, c_f]),
    "b2_sector_table_host": (c_int, [c_int, C.POINTER(C.c_int32)]),
    "b2_sector_table_rotated_host": (c_int, [c_int, C.c_double, C.POINTER(C.c_int32)]),
    "b2_sector_map_i32_ex": (c_int, [c_f, C.POINTER(SectorFrame), c_f, c_i64, c_i64, c_i64, c_int, c_f]),
    "b2_strain_sector_fwd_ex": (c_int, [c_f, c_f, c_f, C.POINTER(SectorFrame), c_f, c_f, c_i64, c_i64, c_i64, c_i64,
                                        c_int, c_int, c_f]),
    "b2_strain_sector_bwd_ex": (c_int, [c_f, c_f, c_f, c_f, C.POINTER(SectorFrame), c_f, c_f, c_i64, c_i64, c_i64,
                                        c_i64, c_int, c_int, c_f]),
    "b2_shoot_workspace_bytes_flags": (c_i64, [c_i64, c_i64, c_i64, c_i64, c_int, c_int]),
    "b2_shoot_bwd_workspace_bytes_flags": (c_i64, [c_i64, c_i64, c_i64, c_int]),
    "b2_shoot_bwd_ex": (c_int, [C.POINTER(ShootBwdArgs), c_f, c_i64, c_f]),
    "b2_sizeof_shoot_bwd_args": (c_i64, []),
    "b2_mask_moments": (c_int, [c_f, c_f, c_i64, c_i64, c_i64, c_f]),
    "b2_sector_map_i32": (c_int, [c_f, c_f, c_f, c_i64, c_i64, c_i64, c_int, c_f]),
    "b2_strain_sector_fwd": (c_int, [c_f, c_f, c_f, c_f, c_f, c_f, c_i64, c_i64, c_i64, c_i64, c_int, c_int, c_f]),
    "b2_strain_sector_bwd": (c_int, [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_i64, c_i64, c_i64, c_i64, c_int, c_int, c_f]),
    "b2_shoot_workspace_bytes": (c_i64, [c_i64, c_i64, c_i64, c_i64, c_int]),
    "b2_shoot_fwd": (c_int, [C.POINTER(ShootArgs), c_f, c_i64, c_f]),
    "b2_shoot_bwd_workspace_bytes": (c_i64, [c_i64, c_i64, c_i64]),
    "b2_shoot_bwd": (c_int, [c_f, c_f, c_f, c_f, c_f, c_f, c_i64, c_i64, c_i64, c_int, c_float, c_float, c_float,
                             c_float, c_int, c_int, c_f, c_i64, c_f]),
    "b2_sizeof_shoot_args": (c_i64, []),
    "b2_shoot_bwd_loss": (c_int, [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_i64, c_i64, c_i64, c_int, c_float, c_float,
                                  c_float, c_float, c_int, c_int, c_f, c_i64, c_f]),
    "b2_recon_loss_terms": (c_int, [c_f, c_f, c_f, c_f, c_f, c_i64, c_i64, c_i64, c_f]),
    "b2_warp_sqerr_bwd": (c_int, [c_f, c_f, c_f, c_f, c_f, c_f, c_i64, c_i64, c_i64, c_i64, c_int, c_i64, c_i64,
                                  c_int, c_int, c_f]),
    "b2_augment_volume": (c_int, [c_f, c_f, c_f, c_f, c_i64, c_i64, c_i64, c_i64, c_f]),
    "b2_roll_rows": (c_int, [c_f, c_f, c_f, c_i64, c_i64, c_i64, c_f]),
    "b2_regroup_pairs": (c_int, [c_f, c_f, c_f, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_f]),
    "b2_regroup_pairs_bwd": (c_int, [c_f, c_f, c_f, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_f]),
    "b2_pack_binary_u8_host": (c_int, [c_f, c_f, c_i64, c_int]),
    "b2_unpack_u8": (c_int, [c_f, c_f, c_i64, c_f]),
    "b2_unpack_bits": (c_int, [c_f, c_f, c_i64, c_f]),
    "b2_device_sm_count": (c_int, [c_int]),
    "b2_shoot_cluster_occupancy": (c_int, [c_f, c_f]),
}


def lib_path() -> pathlib.Path:
    return _LIB_PATH


def lib():
    """Load (once) and return the C-ABI library; raise loudly if it is missing."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(
                f"CUDA extension {_LIB_PATH} is missing - build it with "
                "`python __graft_entry__.py` (nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(str(_LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        for what, have, want in (("b2_shoot_args", L.b2_sizeof_shoot_args(), C.sizeof(ShootArgs)),
                                 ("b2_shoot_bwd_args", L.b2_sizeof_shoot_bwd_args(), C.sizeof(ShootBwdArgs))):
            if have != want:
                raise RuntimeError(f"{_LIB_PATH}: {what} is {have} bytes in the library, {want} in the binding - "
                                   "rebuild with `python __graft_entry__.py`")
        _lib = L
    return _lib


class B2Error(RuntimeError):
    pass


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = lib().b2_error_string(int(code)).decode()
        raise B2Error(f"{what or 'b2lddmm'} failed with code {code}: {msg}")


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream(device=None):
    """The caller's current stream ON ``device`` (default: the current device) as a ``cudaStream_t``."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def device_guard(fn):
    """Run ``fn`` with the device of its first CUDA tensor argument current, so that ``stream()`` and every
    launch inside target the tensors' device even when another device is current (multi-GPU processes)."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kw):
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kw)
                break
        return fn(*args, **kw)
    return wrapped


class on_device:
    """``with on_device(t_or_dev):`` makes the tensors' device current for the launches inside, so a call on
    cuda:1 tensors while cuda:0 is current launches on cuda:1's context and ITS current stream."""

    def __init__(self, dev):
        if isinstance(dev, torch.Tensor):
            dev = dev.device
        self._g = torch.cuda.device(dev)

    def __enter__(self):
        self._g.__enter__()
        return self

    def __exit__(self, *exc):
        return self._g.__exit__(*exc)


def require_cuda(*tensors, dtype=torch.float32):  # noqa: C901
    """All tensors must be contiguous CUDA tensors of one device (no CPU fallback).  Returns that device."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise B2Error("b2lddmm ops run on CUDA tensors only (there is no CPU fallback)")
        if dtype is not None and t.dtype != dtype:
            raise B2Error(f"expected dtype {dtype}, got {t.dtype}")
        if not t.is_contiguous():
            raise B2Error("expected a contiguous tensor")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise B2Error("all tensors must live on the same device")
    return dev


_launches = 0


def count_launch(n: int = 1) -> None:
    """Bookkeeping for bench.py's ``gpu_launches``: kernels of OUR library launched."""
    global _launches
    _launches += n


def launches() -> int:
    return _launches

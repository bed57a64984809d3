"""ctypes wrapper of the plain-C oracle (oracle/lddmm_c.c) - TEST INFRASTRUCTURE ONLY.

``build()`` compiles it with ``gcc -O3 -fopenmp`` into ``oracle/_build/liboracle_c.so`` (git-ignored, travels to the
GPU box with the snapshot); ``__graft_entry__.build()`` calls it.  ``forward_volume`` mirrors
:func:`oracle.path.forward_volume` for the Lagrangian split and is the multi-threaded CPU baseline of bench.py.
"""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
SRC = HERE / "lddmm_c.c"
LIB = HERE / "_build" / "liboracle_c.so"
_lib = None


def build(force: bool = False) -> pathlib.Path:
    LIB.parent.mkdir(parents=True, exist_ok=True)
    if force or not LIB.exists() or LIB.stat().st_mtime < SRC.stat().st_mtime:
        cmd = ["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", "-std=gnu11", str(SRC), "-o", str(LIB), "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"gcc failed for {SRC.name}:\n{r.stdout}\n{r.stderr}")
    return LIB


def available() -> bool:
    return LIB.exists()


def lib():
    global _lib
    if _lib is None:
        if not LIB.exists():
            build()
        L = C.CDLL(str(LIB))
        L.b2o_forward_volume.restype = C.c_int
        L.b2o_forward_volume.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_float] * 3 + [C.c_int] * 2 \
            + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 2
        L.b2o_sector_table.restype = None
        L.b2o_sector_table.argtypes = [C.c_int, C.c_void_p]
        L.b2o_max_threads.restype = C.c_int
        _lib = L
    return _lib


def max_threads() -> int:
    return int(lib().b2o_max_threads())


def forward_volume(v0: torch.Tensor, vol: torch.Tensor, params=(1.0, 0.1, 0.05), num_steps: int = 10,
                   n_sectors: int = 126, n_frames: int = 40, nthreads: int = 0, theta0=None, clockwise=None):
    """v0 (B*(T-1),2,H,W), vol (B,1,T,H,W): fp32 CPU tensors.  Returns the forward_volume dict (Lagrangian split).
    ``theta0`` (radians) / ``clockwise``: per-slice sector frame (B entries or scalars; default 0 / True).
    ``nthreads`` > 0 sets the OpenMP team size explicitly (0 = OpenMP default, which OMP_NUM_THREADS caps)."""
    B, one, T, H, W = vol.shape
    assert one == 1 and tuple(v0.shape) == (B * (T - 1), 2, H, W)
    v0n = np.ascontiguousarray(v0.detach().cpu().numpy(), dtype=np.float32)
    voln = np.ascontiguousarray(vol.detach().cpu().numpy(), dtype=np.float32)
    P, T1 = B * (T - 1), T - 1
    m0 = np.empty((P, 2, H, W), np.float32)
    vel = np.empty_like(m0)
    u = np.empty_like(m0)
    sdef = np.empty((P, 1, H, W), np.float32)
    S = np.empty((B, 1, n_sectors, n_frames), np.float32)
    th = cw = None
    if theta0 is not None:
        th = np.ascontiguousarray(np.broadcast_to(np.asarray(theta0, np.float64).reshape(-1), (B,)))
    if clockwise is not None:
        cw = np.ascontiguousarray(np.broadcast_to(np.asarray(clockwise).reshape(-1) != 0, (B,)).astype(np.int32))
    rc = lib().b2o_forward_volume(v0n.ctypes.data, voln.ctypes.data, B, T, H, W, int(num_steps),
                                  float(params[0]), float(params[1]), float(params[2]), int(n_sectors), int(n_frames),
                                  m0.ctypes.data, vel.ctypes.data, u.ctypes.data, sdef.ctypes.data, S.ctypes.data,
                                  int(nthreads), th.ctypes.data if th is not None else None,
                                  cw.ctypes.data if cw is not None else None)
    if rc != 0:
        raise RuntimeError(f"b2o_forward_volume failed ({rc})")
    t = torch.from_numpy
    return {"strain_matrix": t(S), "deformed_source": t(sdef).view(B, 1, T1, H, W), "velocity": t(vel),
            "momentum": t(m0), "displacement": t(u)}

"""CPU tests of the host side: C-ABI symbols, error behaviour, sharding, gloo all-reduce."""
import ctypes
import os
import pathlib
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_header_symbols_exported(pkg):
    """The shared library loads and exports every entry point include/b2lddmm.h declares."""
    hdr = (ROOT / "include" / "b2lddmm.h").read_text()
    declared = set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"b2_shoot_args", "b2_shoot_bwd_args", "b2_sector_frame"}
    assert len(declared) >= 25
    L = ctypes.CDLL(str(pkg._lib.lib_path()))
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(pkg._lib.SIGNATURES) == declared
    assert pkg._lib.lib().b2_version() >= 100


def test_sector_table_matches_oracle(pkg, oracle):
    for n in (3, 18, 126, 256):
        buf = (ctypes.c_int32 * (2 * n))()
        assert pkg._lib.lib().b2_sector_table_host(n, buf) == 0
        assert np.array_equal(np.array(list(buf)).reshape(n, 2), oracle.sector_boundaries(n))
    assert pkg._lib.lib().b2_sector_table_host(2, (ctypes.c_int32 * 4)()) == -5
    assert pkg._lib.lib().b2_sector_table_host(126, None) == -1
    # rotated tables (per-slice sector frame): same integers as the oracle's for arbitrary start angles
    for th in (0.0, 1.09432890732119, -1.878849107818673, 6.0, -0.3):
        buf = (ctypes.c_int32 * (2 * 126))()
        assert pkg._lib.lib().b2_sector_table_rotated_host(126, th, buf) == 0
        assert np.array_equal(np.array(list(buf)).reshape(126, 2), oracle.sector_boundaries(126, th))
    assert pkg._lib.lib().b2_sector_table_rotated_host(126, float("nan"), (ctypes.c_int32 * 252)()) == -5


def test_struct_sizes_match_header(pkg):
    L = pkg._lib.lib()
    assert L.b2_sizeof_shoot_args() == ctypes.sizeof(pkg._lib.ShootArgs)
    assert L.b2_sizeof_shoot_bwd_args() == ctypes.sizeof(pkg._lib.ShootBwdArgs)
    # the adjoint workspace is sized per path: fused (resident CTAs x 3 fields, as much again for the tail pairs of the
    # dynamic schedule, ticket + flags) far below op-level (5 P fields)
    fused = L.b2_shoot_bwd_workspace_bytes_flags(1536, 128, 128, 0)
    oplevel = L.b2_shoot_bwd_workspace_bytes_flags(1536, 128, 128, pkg._lib.FLAG_OPLEVEL)
    assert 0 < fused < oplevel / 8 and oplevel >= 5 * 1536 * 2 * 128 * 128 * 4
    assert fused >= 2 * 148 * 3 * 2 * 128 * 128 * 4
    assert L.b2_shoot_bwd_workspace_bytes(1536, 128, 128) == fused
    assert L.b2_shoot_bwd_workspace_bytes_flags(4, 100, 100, 0) > 0     # op-level sized; rejected at call time (FFT size)


def test_argument_errors_without_gpu(pkg):
    """Invalid arguments are rejected before any CUDA work (negative codes, never a throw)."""
    L = pkg._lib.lib()
    assert L.b2_interp_fwd(None, None, None, 1, 1, 1, 1, 8, 8, 1.0, 0, None) == -1
    assert L.b2_fluid_workspace_bytes(4, 128, 128) == 0
    assert L.b2_fluid_workspace_bytes(4, 256, 256) == 4 * 256 * 256 * 8
    assert L.b2_shoot_workspace_bytes(1, 1, 100, 100, 10) > 0  # path-B sized; rejected at call time
    assert b"FFT" in L.b2_error_string(-4)
    with pytest.raises(RuntimeError):
        pkg.interp(torch.zeros(1, 1, 8, 8), torch.zeros(1, 2, 8, 8))      # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        pkg.FluidMetric((1.0, 0.1, 0.0))
    with pytest.raises(NotImplementedError):
        pkg.build_model({"type": "Nope"})


def test_idle_sm_split_planner(pkg):
    """Host logic of the 256x256 idle-SM split (shooting._idle_split_slices), no GPU: with 33 co-resident clusters
    and 16 stranded SMs (a 148-SM B200) it hands over the slice counts measured best on the device (DESIGN.md
    section 6), nothing when no SM is stranded or there is a single slice, and never more than half the batch."""
    import torch
    sh = pkg.shooting
    dev = torch.device("cuda", 0)
    saved = dict(sh._cluster_occ)
    try:
        sh._cluster_occ[0] = (33, 16)
        assert sh._idle_split_slices(16, 49, dev) == 2
        assert sh._idle_split_slices(32, 49, dev) == 3
        assert 26 <= sh._idle_split_slices(256, 49, dev) <= 30
        assert sh._idle_split_slices(1, 49, dev) == 0
        for B in range(2, 40):
            assert 0 <= sh._idle_split_slices(B, 24, dev) <= B // 2
        # pair granularity: whole cluster rounds on the cluster arm, the op-level arm ends inside them, and the cut
        # the sweep on the device found best (58 of 784 pairs, 149 of 1568)
        import math
        assert sh._idle_split_pairs(16, 49, dev, 2) == 58 and sh._idle_split_pairs(32, 49, dev, 3) == 149
        for B in (16, 32, 64, 256):
            b2 = sh._idle_split_slices(B, 49, dev)
            p2 = sh._idle_split_pairs(B, 49, dev, b2)
            rounds = (B * 49 - p2) // 33
            assert 0 < p2 <= B * 49 // 2 and (B * 49 - p2) % 33 == 0
            assert 0.27 * p2 <= rounds < math.ceil(B * 49 / 33)
        sh.idle_sm_pair_granular = False
        assert sh._idle_split_pairs(16, 49, dev, 2) == 98
        sh.idle_sm_pair_granular = True
        sh._cluster_occ[0] = (37, 0)              # a device whose GPCs pack whole clusters: nothing to gain
        assert sh._idle_split_slices(64, 49, dev) == 0
        sh._cluster_occ[0] = (0, 0)               # clusters unavailable
        assert sh._idle_split_slices(64, 49, dev) == 0
        sh._cluster_occ[0] = (33, 16)
        sh.idle_sm_split = False
        assert sh._idle_split_slices(64, 49, dev) == 0
    finally:
        sh.idle_sm_split = True
        sh._cluster_occ.clear()
        sh._cluster_occ.update(saved)


def test_shard_slices(pkg):
    for n in (1, 7, 64, 256):
        for ws in (1, 2, 3, 8):
            spans = [pkg.parallel.shard_slices(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pkg.parallel.shard_slices(4, 2, 2)


def test_models_interface(pkg):
    """Module-level contract of main.py:42-46 / trainer :59,183,243 (CPU-constructible, CUDA to run)."""
    import copy
    j = pkg.build_model({"type": "JointRegisterStrainMatNet", "n_strain_matrix_frames": 40})
    l = pkg.build_model({"type": "NetStrainMat2LMA", "num_conv_layers": 3, "inner_conv_channel_num": 16,
                         "input_channel_num": 1, "n_frames": 40, "n_sectors": 126, "n_classes": 1})
    assert j.sigma == 0.03 and len(list(j.parameters())) > 0 and copy.deepcopy(j).state_dict().keys() == j.state_dict().keys()
    out = l(torch.zeros(3, 1, 126, 40))
    assert out["TOS"].shape == (3, 126)
    S = torch.randn(2, 1, 126, 40)
    assert torch.linalg.matrix_rank(pkg.models.svd_smooth(S, 5)[0, 0]) == 5


def test_svd_smooth_matches_reference_svddenoise(pkg):
    """``svd_smooth`` vs the output of the reference's own ``SVDDenoise`` (DENSE_utils.py:11-14; golden made by
    tests/golden/make_golden.py), both gradient modes; the projection gradient is U_r U_r^T g."""
    g = np.load(ROOT / "tests" / "golden" / "ref_sectors.npz")
    S = torch.from_numpy(g["svd_in"])                                    # float64, as the reference computes it
    for rank in (3, 5):
        want = torch.from_numpy(g[f"svd_rank{rank}"])
        for mode in ("projection", "exact"):
            got = pkg.models.svd_smooth(S[None, None], rank, grad=mode)[0, 0]
            assert (got - want).abs().max() < 1e-12 * want.abs().max(), (rank, mode)
        got32 = pkg.models.svd_smooth(S[None, None].float(), rank)[0, 0]
        assert (got32.double() - want).abs().max() < 2e-5 * want.abs().max()      # fp32 SVD
    x = S[None, None].clone().requires_grad_(True)
    gout = torch.randn(1, 1, 126, 40, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    pkg.models.svd_smooth(x, 5).backward(gout)
    U = torch.linalg.svd(S, full_matrices=False)[0][:, :5]
    assert (x.grad[0, 0] - U @ (U.T @ gout[0, 0])).abs().max() < 1e-12
    # rank-deficient input (edge-padded strain matrix): the projection gradient stays finite
    pad = torch.cat([S[:, :20], S[:, 19:20].expand(126, 20)], dim=1)[None, None].clone().requires_grad_(True)
    pkg.models.svd_smooth(pad, 5).sum().backward()
    assert torch.isfinite(pad.grad).all()
    with pytest.raises(ValueError):
        pkg.models.svd_smooth(S[None, None], 5, grad="nope")


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
import __graft_entry__ as g
pkg = g.load_package()
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, ws = dist.get_rank(), dist.get_world_size()
torch.manual_seed(0)
net = pkg.build_model({{"type": "NetStrainMat2LMA"}})
a, b = pkg.parallel.shard_slices(6, rank, ws)
x = torch.randn(6, 1, 126, 40, generator=torch.Generator().manual_seed(1))
y = torch.randn(6, 126, generator=torch.Generator().manual_seed(2))
loss = ((net(x[a:b])["TOS"] - y[a:b]) ** 2).sum() / 6 * ws   # so that the rank-average equals the full-batch grad
loss.backward()
n = pkg.parallel.allreduce_gradients(list(net.parameters()))
# hook-driven reducer (all-reduce launched from inside backward, two buckets): same averaged gradient
net2 = pkg.build_model({{"type": "NetStrainMat2LMA"}})
net2.load_state_dict(net.state_dict())
ps = list(net2.parameters())
red = pkg.parallel.GradientAllReducer([ps[len(ps) // 2:], ps[:len(ps) // 2]])
(((net2(x[a:b])["TOS"] - y[a:b]) ** 2).sum() / 6 * ws).backward()
n2 = red.finish()
err2 = max((p.grad - q.grad).abs().max().item() for p, q in zip(net.parameters(), net2.parameters()))
assert n2 == 2 and err2 < 1e-6, (n2, err2)
ref = pkg.build_model({{"type": "NetStrainMat2LMA"}})
ref.load_state_dict(net.state_dict())
(((ref(x)["TOS"] - y) ** 2).sum() / 6).backward()
err = max((p.grad - q.grad).abs().max().item() for p, q in zip(net.parameters(), ref.parameters()))
S = pkg.parallel.gather_strain_matrices(x[a:b], 6)
ok = (rank != 0) or torch.equal(S, x)
print(f"rank{{rank}} ncoll={{n}} err={{err:.2e}} gather_ok={{ok}}")
assert n == 1 and err < 1e-5 and ok
dist.destroy_process_group()
"""


def test_gloo_world2_gradient_allreduce(tmp_path):
    """N>1 host logic on CPU: slice sharding + bucketed gradient all-reduce == full-batch gradient."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=str(ROOT)))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29613")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)


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
    declared -= {"b2_shoot_args"}
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


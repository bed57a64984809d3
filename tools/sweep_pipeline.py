"""Sweep the host pipeline's chunk size and staging depth at configs[1] (e2e ms per step, streaming submit/get).

    python tools/sweep_pipeline.py [bits|u8]
"""
import json
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402
import bench  # noqa: E402


def main():
    fmt = sys.argv[1] if len(sys.argv) > 1 else "bits"
    pkg = g.load_package()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, T, H, W = bench.B_PER_GPU, bench.T_FRAMES, bench.H, bench.W
    metric = pkg.FluidMetric(bench.PARAMS)
    vol_h, v0_h = bench.make_inputs(pkg, B, seed=2434)
    v0_h = v0_h.pin_memory()
    if fmt == "bits":
        m_h = torch.from_numpy(np.packbits(vol_h.numpy() > 0.5, axis=-1)).pin_memory()
    else:
        m_h = vol_h.to(torch.uint8).pin_memory()
    spin = torch.randn(4096, 4096, device=dev)
    for _ in range(200):
        spin = (spin @ spin).clamp_(-1, 1)
    torch.cuda.synchronize()
    rows = []
    for rep in range(2):
        for cs in (8, 11, 13, 16, 22, 32):
            for ns in (3, 4, 6):
                pipe = pkg.HostPipeline(B, T, H, W, metric, num_steps=bench.S_STEPS, n_sectors=bench.N_SECTORS,
                                        n_frames=bench.N_FRAMES, device=dev, chunk_slices=cs, n_stages=ns)
                pending = []

                def step():
                    pending.append(pipe.submit(v0_h, m_h))
                    if len(pending) > 1:
                        pending.pop(0).get()
                for _ in range(5):
                    step()
                while pending:
                    pending.pop(0).get()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(30):
                    step()
                while pending:
                    pending.pop(0).get()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 30
                rows.append({"rep": rep, "chunk_slices": cs, "n_stages": ns, "ms": ms})
                print(json.dumps(rows[-1]), flush=True)
                del pipe


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Op-level kernel timings (CUDA events) with achieved DRAM-contract bandwidth: interp / compose / warp forward.

  python tools/bench_ops.py            # current library
  B2LDDMM_LIB=build/variants/<name>/libb2lddmm.so python tools/bench_ops.py      # an A/B build

Contract bytes per pixel (fp32): interp (2 + 2C) x 4 (u in, I in, out), compose 6 x 4, warp of a shared source
(2 + C) x 4 + source once per slice.
"""
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402


def timed(fn, steps=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    pkg = g.load_package()
    dev = torch.device("cuda:0")
    peak = 6545.9
    try:
        peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        pass
    for (P, H, W) in ((1536, 128, 128), (784, 256, 256)):
        N = H * W
        u = 3.0 * torch.randn(P, 2, H, W, device=dev)
        u = pkg.FluidMetric((1.0, 0.1, 0.05)).sharp(u)
        u = u * (3.0 / u.abs().amax())
        for C in (1, 2):
            I = torch.randn(P, C, H, W, device=dev)
            with torch.no_grad():
                ms = timed(lambda: pkg.interp(I, u, 1.0))
            gb = P * N * (2 + 2 * C) * 4 / 1e9
            print(json.dumps({"op": f"interp C={C}", "P": P, "grid": [H, W], "us": 1e3 * ms, "GBps": gb / (ms * 1e-3),
                              "frac_of_peak": gb / (ms * 1e-3) / peak}))
        v = torch.randn(P, 2, H, W, device=dev)
        with torch.no_grad():
            ms = timed(lambda: pkg.compose_disp_vel(u, v, -0.1))
        gb = P * N * 6 * 4 / 1e9
        print(json.dumps({"op": "compose_disp_vel", "P": P, "grid": [H, W], "us": 1e3 * ms, "GBps": gb / (ms * 1e-3),
                          "frac_of_peak": gb / (ms * 1e-3) / peak}))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Secondary measurements for the other BASELINE.json configs (not the driver's bench line).

  python tools/bench_configs.py c3      # configs[2]: C2 batch + EPDiff adjoint backward (training-mode gradients)
  python tools/bench_configs.py c3f     # same with the fused loss epilogue (per-pair loss terms from the kernel)
  python tools/bench_configs.py c4      # configs[3]: 256x256, 50 frames (4-CTA cluster kernel; B2_NO_CLUSTER=1 = op-level path), a shard of 16 slices

Prints one JSON line per config with pairs/s and the op-level roofline fraction.
"""
import json
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402

PARAMS = (1.0, 0.1, 0.05)


def timed(fn, steps, warmup):
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
    which = sys.argv[1] if len(sys.argv) > 1 else "c3"
    pkg = g.load_package()
    dev = torch.device("cuda:0")
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    metric = pkg.FluidMetric(PARAMS)
    S = 10
    if which in ("c3", "c3f"):
        B, T, H, W = 64, 25, 128, 128
    else:
        B, T, H, W = 16, 50, 256, 256
    T1, P, N = T - 1, B * (T - 1), H * W
    vol = pkg.synthetic.synthetic_masks(B, T, H, W).to(dev)
    v0 = pkg.synthetic.synthetic_v0(P, H, W, seed=5, max_disp=3.0).to(dev)
    src_vol, tar_vol = pkg.data.split_vol_to_registration_pairs(vol, "Lagrangian", 3)
    bytes_fwd = 4 * (15 + 16 * S) * N
    if which in ("c3", "c3f"):
        Sgt = torch.zeros(B, 1, 126, 40, device=dev)
        vg = v0.clone().requires_grad_(True)
        fused = which == "c3f"            # loss epilogue: per-pair sums from the kernel, seedless adjoints
        crit = pkg.RegistrationReconstructionLoss(0.03, 0.1)

        def step():
            vg.grad = None
            out = pkg.shoot_warp_strain(vg, src_vol, tar_vol, metric, num_steps=S, loss_terms=fused)
            loss = crit(out, {"registration_target": tar_vol}) \
                + 1000.0 * torch.mean((out["strain_matrix"] - Sgt) ** 2)
            loss.backward()

        ms = timed(step, 5, 2)
        nbytes = 3 * bytes_fwd
        name = "configs[2]: C2 batch + EPDiff adjoint backward" + (" (fused loss epilogue)" if fused else "")
    else:
        def step():
            with torch.no_grad():
                pkg.shoot_warp_strain(v0, src_vol, tar_vol, metric, num_steps=S)

        ms = timed(step, 3, 1)
        nbytes = bytes_fwd
        name = f"configs[3] shard: {B} slices x {T} frames 256x256 forward ({'op-level path' if os.environ.get('B2_NO_CLUSTER') else 'cluster kernel'})"
    ach = P * nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"config": name, "pairs": P, "ms_per_step": ms, "pairs_per_s": P / (ms * 1e-3),
                      "roofline": {"achieved_GBps": ach, "peak_GBps": peak, "frac": ach / peak,
                                   "algorithmic_bytes_per_pair": nbytes}}))


if __name__ == "__main__":
    main()

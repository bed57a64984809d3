#!/usr/bin/env python
"""Calibrate `shooting.idle_sm_pair_cost`: time the 256x256 inference batch with b2 = 0, 1, 2, ... trailing slices on
the op-level path (second stream, the SMs the 4-CTA clusters strand) and print ms per step for each.

  python tools/sweep_idle_split.py [slices] [b2 values, comma separated] [P2 values (trailing PAIRS), comma separated]
"""
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import torch  # noqa: E402
import bench_configs as bc  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    b2s = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3]
    pkg, dev, _, _ = bc._init()
    sh = pkg.shooting
    auto = sh._idle_split_slices
    metric = pkg.FluidMetric(bc.PARAMS)
    T = 50
    vol = pkg.synthetic.synthetic_masks(B, T, 256, 256).to(dev)
    v0 = bc.device_v0(pkg, B * (T - 1), 256, 256, 5, dev, chunk=784)
    src_vol, tar_vol = pkg.data.split_vol_to_registration_pairs(vol, "Lagrangian", 3)

    def step():
        with torch.no_grad():
            pkg.shoot_warp_strain(v0, src_vol, tar_vol, metric, num_steps=bc.S_STEPS)

    res = {}
    for b2 in b2s:
        sh._idle_split_slices = (lambda n: (lambda B_, T1_, dev_: n))(b2)
        res[str(b2)] = round(bc.timed(step, 3, 1), 3)
    res_p = {}
    auto_p = sh._idle_split_pairs
    if len(sys.argv) > 3:                      # cuts at pair granularity (pair ranges of b2_shoot_args)
        sh._idle_split_slices = lambda B_, T1_, dev_: 1
        for p2 in [int(x) for x in sys.argv[3].split(",")]:
            sh._idle_split_pairs = (lambda n: (lambda B_, T1_, dev_, b2_: n))(p2)
            res_p[str(p2)] = round(bc.timed(step, 3, 1), 3)
    sh._idle_split_slices, sh._idle_split_pairs = auto, auto_p
    res["auto"] = round(bc.timed(step, 3, 1), 3)
    b2 = auto(B, T - 1, dev)
    print(json.dumps({"slices": B, "pairs": B * (T - 1), "ms_by_b2": res, "ms_by_p2": res_p, "auto_b2": b2,
                      "auto_p2": auto_p(B, T - 1, dev, b2) if b2 else 0}))


if __name__ == "__main__":
    main()

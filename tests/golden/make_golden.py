"""Generate golden vectors from the reference's own importable code.

Run in the authoring container only (needs /root/reference, which does not
exist on the GPU box):

    python tests/golden/make_golden.py

The reference tree cannot run its hot path (missing ``models`` package and
``lagomorph``), so the only reference-generated goldens
are for the callers either side of it:

* ``split_vol_to_registration_pairs``  /root/reference/modules/data/__init__.py:93-121
* ``align_n_frames_to``                /root/reference/modules/data/datareader/DENSE_IO_utils.py:2-46
* ``LossCalculator`` with the shipped loss config
                                       /root/reference/modules/loss/loss_calculator.py:104-126,
                                       /root/reference/configs/config.json:164-196
* the rotation <-> sector-roll convention
                                       /root/reference/modules/data/augmentation/affine.py:52-87
  (recorded as the constants it pins; skimage itself is not installed here).

Outputs ``tests/golden/ref_boundary.npz`` (committed).
"""
import importlib.util
import json
import pathlib
import sys
import types

import numpy as np
import torch

REF = pathlib.Path("/root/reference")
OUT = pathlib.Path(__file__).resolve().parent / "ref_boundary.npz"


def _import_reference():
    # skimage is absent in this image; the functions we need never call it.
    sk = types.ModuleType("skimage")
    skt = types.ModuleType("skimage.transform")
    skt.rotate = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("skimage stub"))
    skm = types.ModuleType("skimage.morphology")
    skm.dilation = skt.rotate
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.transform", skt)
    sys.modules.setdefault("skimage.morphology", skm)
    sys.path.insert(0, str(REF))
    import modules.data as refdata  # noqa: E402
    from modules.loss.loss_calculator import LossCalculator  # noqa: E402
    spec = importlib.util.spec_from_file_location(
        "ref_dense_io_utils", REF / "modules/data/datareader/DENSE_IO_utils.py")
    io_utils = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(io_utils)
    return refdata, io_utils, LossCalculator


def main():
    refdata, io_utils, LossCalculator = _import_reference()
    g = torch.Generator().manual_seed(2434)   # /root/reference/configs/config.json:129
    out = {}

    vol = torch.rand(2, 1, 5, 6, 7, generator=g)
    out["split_vol"] = vol.numpy()
    for method in ("Lagrangian", "Eulerian"):
        for od in (2, 3):
            s, t = refdata.split_vol_to_registration_pairs(vol, split_method=method, output_dim=od)
            out[f"split_{method}_{od}_src"] = s.numpy()
            out[f"split_{method}_{od}_tar"] = t.numpy()

    a = torch.rand(4, 5, 7, generator=g).numpy()
    out["align_in"] = a
    out["align_crop3"] = io_utils.align_n_frames_to(a, 3)
    out["align_pad10"] = io_utils.align_n_frames_to(a, 10)
    out["align_pad9_axis1"] = io_utils.align_n_frames_to(a, 9, frame_idx=1)
    out["align_crop2_axis0"] = io_utils.align_n_frames_to(a, 2, frame_idx=0)

    cfg = json.loads((REF / "configs/config.json").read_text())
    calc = LossCalculator(cfg["losses"])
    B, T1, H, W = 2, 3, 8, 8
    pred = {
        "strainmat": torch.randn(B, 1, 126, 40, generator=g) * 0.1,
        "deformed_source": torch.rand(B, 1, T1, H, W, generator=g),
        "TOS": torch.rand(B, 126, generator=g) * 60,
        "velocity": torch.randn(B * T1, 2, H, W, generator=g),
        "momentum": torch.randn(B * T1, 2, H, W, generator=g),
    }
    target = {
        "strainmat": torch.randn(B, 1, 126, 40, generator=g) * 0.1,
        "registration_target": (torch.rand(B, 1, T1, H, W, generator=g) > 0.5).float(),
        "TOS": torch.rand(B, 126, generator=g) * 60,
    }
    total, parts = calc(pred, target)
    for k, v in pred.items():
        out[f"loss_pred_{k}"] = v.numpy()
    for k, v in target.items():
        out[f"loss_target_{k}"] = v.numpy()
    out["loss_total"] = np.float64(total.item())
    for k, v in parts.items():
        out[f"loss_part_{k}"] = np.float64(v)
    out["loss_sigma"] = np.float64(cfg["losses"]["registration_reconstruction"]["sigma"])
    out["loss_reg_weight"] = np.float64(cfg["losses"]["registration_reconstruction"]["regularization_weight"])
    out["loss_weights"] = np.array([cfg["losses"][k]["weight"] for k in
                                    ("registration_reconstruction", "registration_supervision",
                                     "TOS_regression")], dtype=np.float64)

    # constants pinned by the reference around the path
    net = cfg["networks"]
    out["n_sectors"] = np.int64(net["LMA"]["n_sectors"])
    out["n_strain_frames"] = np.int64(net["joint_register_strainmat"]["n_strain_matrix_frames"])
    out["seed"] = np.int64(cfg["training"]["seed"])

    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes, {len(out)} arrays)")


if __name__ == "__main__":
    main()

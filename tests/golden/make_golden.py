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

* ``translate`` / ``rotate``           /root/reference/modules/data/augmentation/affine.py:24-87
  (numpy parts as is; the skimage call is recorded through a stub: angle and keyword arguments).

* ``merge_data_of_same_slice_from_batch``
                                       /root/reference/modules/trainer/joint_registration_regression_trainer.py:54-120

* ``spl2patchSA`` (the 126-sector polar mesh: start angle ``theta0 = arctan2(PositionB - PositionA)``, direction by
  the ``Clockwise`` flag, 18 x 7 = 126 angular samples, mid-wall = layer 3) and ``SVDDenoise``
                                       /root/reference/modules/data/utils/DENSE_utils.py:11-14,177-295
  (PyQt5 is absent: stubbed, it is only used by an unrelated screen-size helper).

Outputs ``tests/golden/ref_boundary.npz``, ``ref_augment.npz``, ``ref_regroup.npz`` and ``ref_sectors.npz`` (committed).
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
OUT_AUG = pathlib.Path(__file__).resolve().parent / "ref_augment.npz"
OUT_REGROUP = pathlib.Path(__file__).resolve().parent / "ref_regroup.npz"
OUT_SECTORS = pathlib.Path(__file__).resolve().parent / "ref_sectors.npz"
SKROTATE_CALLS = []


def _import_reference():
    # skimage is absent in this image; the functions we need never call it.
    sk = types.ModuleType("skimage")
    skt = types.ModuleType("skimage.transform")
    def _skrotate_stub(image, angle, **kw):
        # skimage is not installed: record how the reference calls it and hand the image back unchanged
        SKROTATE_CALLS.append((float(angle), dict(kw)))
        return image
    skt.rotate = _skrotate_stub
    skm = types.ModuleType("skimage.morphology")
    skm.dilation = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("skimage stub"))
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
    make_augment_golden()
    make_regroup_golden()
    make_sector_golden()


def make_sector_golden():
    """The reference's own sector mesh and SVD smoothing (DENSE_utils.py:177-295, :11-14) on synthetic contours.

    ``spl2patchSA`` is run on two concentric circular contours (epicardium r = 110, endocardium r = 90) about
    ``PositionA``; the centres of its 126 mid-wall faces (``layerid == 3``, DENSE_utils.py:323), in mesh order, are
    the golden: face k must be classified as sector k by the oracle and by the CUDA classifier in the frame
    (theta0, clockwise) of that case.  Contour points are (x, y) = (column, row) image coordinates."""
    qt = types.ModuleType("PyQt5")
    qt.QtWidgets = types.ModuleType("PyQt5.QtWidgets")
    sys.modules.setdefault("PyQt5", qt)
    sys.modules.setdefault("PyQt5.QtWidgets", qt.QtWidgets)
    if not hasattr(np, "Inf"):
        np.Inf = np.inf                      # DENSE_utils.py:152 uses the alias numpy 2 removed (singular solves only)
    spec = importlib.util.spec_from_file_location("ref_dense_utils", REF / "modules/data/utils/DENSE_utils.py")
    du = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(du)

    ang = np.linspace(0.0, 2.0 * np.pi, 721)[:, None] + 0.00123          # closed polygons, vertices off the spokes
    origin = np.array([128.0, 128.0])                                    # (x, y)
    circle = lambda r: np.hstack([origin[0] + r * np.cos(ang), origin[1] + r * np.sin(ang)])  # noqa: E731
    out = {"origin_xy": origin, "n_cases": np.int64(0)}
    cases = [((200.0, 128.0), 1), ((200.0, 128.0), 0), ((160.0, 190.0), 1), ((160.0, 190.0), 0), ((100.0, 40.0), 1),
             ((100.0, 40.0), 0)]
    for i, (posB, cw) in enumerate(cases):
        datamat = {
            "ROIInfo": types.SimpleNamespace(RestingContour=[circle(110.0), circle(90.0)]),
            "AnalysisInfo": types.SimpleNamespace(PositionA=origin.copy(), PositionB=np.array(posB), Clockwise=cw),
        }
        fv = du.spl2patchSA(datamat)
        faces = fv["faces"][fv["layerid"] == 3] - 1                       # 0-based vertex ids of the mid-wall faces
        assert faces.shape == (126, 4), faces.shape
        centers = fv["vertices"][faces].mean(axis=1)                      # (126, 2) (x, y), mesh order
        out[f"case{i}_posB_xy"] = np.array(posB)
        out[f"case{i}_clockwise"] = np.int64(cw)
        out[f"case{i}_theta0"] = np.float64(np.arctan2(posB[1] - origin[1], posB[0] - origin[0]))   # DENSE_utils.py:198
        out[f"case{i}_midwall_centers_xy"] = centers
        out[f"case{i}_sectorid"] = fv["sectorid"][fv["layerid"] == 3]
    out["n_cases"] = np.int64(len(cases))
    rng = np.random.default_rng(2434)
    mat = rng.standard_normal((126, 40))
    out["svd_in"] = mat
    for rank in (3, 5):
        out[f"svd_rank{rank}"] = du.SVDDenoise(mat.copy(), rank=rank)
    np.savez_compressed(OUT_SECTORS, **out)
    print(f"wrote {OUT_SECTORS} ({OUT_SECTORS.stat().st_size} bytes, {len(out)} arrays)")


def make_regroup_golden():
    """merge_data_of_same_slice_from_batch of joint_registration_regression_trainer.py:54-120.

    The trainer module imports lagomorph / wandb / tensorboard at the top and cannot be imported here, so the
    function definition alone is compiled from the reference file (read in place, nothing copied) and run."""
    import ast
    path = REF / "modules/trainer/joint_registration_regression_trainer.py"
    tree = ast.parse(path.read_text())
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "merge_data_of_same_slice_from_batch")
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[node], type_ignores=[]), str(path), "exec"), ns)
    merge = ns["merge_data_of_same_slice_from_batch"]
    g = torch.Generator().manual_seed(2434)
    ids = ["p1_s0", "p1_s1", "p1_s0", "p2_s0", "p1_s1", "p1_s0", "p1_s0", "p2_s0", "p1_s0"]     # 5 / 2 / 2 pairs
    P, H, W = len(ids), 6, 5
    uniq = list(dict.fromkeys(ids))
    tos_per = {k: torch.rand(126, generator=g) * 60 for k in uniq}
    lab_per = {k: (torch.rand(126, generator=g) > 0.5).long() for k in uniq}
    batch = {
        "slice_full_id": ids,
        "TOS": torch.stack([tos_per[k] for k in ids]),
        "sector_LMA_labels": torch.stack([lab_per[k] for k in ids]),
        "slice_LMA_label": torch.tensor([uniq.index(k) % 2 for k in ids]),
    }
    pred = {"displacement": torch.randn(P, 2, H, W, generator=g)}
    out = {"ids": np.array(ids), "displacement": pred["displacement"].numpy(), "TOS": batch["TOS"].numpy(),
           "sector_LMA_labels": batch["sector_LMA_labels"].numpy(), "slice_LMA_label": batch["slice_LMA_label"].numpy()}
    for F in (3, 4, 7):
        r = merge(batch, pred, F, torch.device("cpu"))
        out[f"F{F}_ids"] = np.array(r["batch_slice_full_ids"])
        out[f"F{F}_fields"] = r["pred_displacement_fields"].numpy()
        out[f"F{F}_TOS"] = r["TOS"].numpy()
        out[f"F{F}_labels"] = r["sector_LMA_labels"].numpy()
        out[f"F{F}_slice_label"] = r["slice_LMA_label"].numpy()
    np.savez_compressed(OUT_REGROUP, **out)
    print(f"wrote {OUT_REGROUP} ({OUT_REGROUP.stat().st_size} bytes, {len(out)} arrays)")


def make_augment_golden():
    """translate / rotate of /root/reference/modules/data/augmentation/affine.py:24-87 on a synthetic datum.

    ``translate`` is pure numpy and runs as is.  ``rotate`` calls skimage (absent): the stub records the angle and
    keyword arguments the reference passes and returns the mask unchanged, so the golden pins the call convention
    and the np.roll of the strain matrix / TOS curve, not skimage's resampling."""
    from modules.data.augmentation import affine as ref_aff  # noqa: E402
    rng = np.random.default_rng(2434)
    H, W, T = 12, 10, 4
    datum = {
        "cine_lv_myo_masks_merged": (rng.random((H, W, T)) > 0.5).astype(np.float32),
        "StrainInfo": {"CCmid": rng.standard_normal((126, 40)).astype(np.float32)},
        "TOSAnalysis": {"TOSfullRes_Jerry": (rng.random(126) * 60).astype(np.float32)},
    }
    out = {"mask": datum["cine_lv_myo_masks_merged"], "strain": datum["StrainInfo"]["CCmid"],
           "tos": datum["TOSAnalysis"]["TOSfullRes_Jerry"]}
    shifts = [(0, 0), (3, -2), (-5, 7), (12, 10), (-13, 1)]
    out["shifts"] = np.array(shifts, dtype=np.int64)
    for i, (ty, tx) in enumerate(shifts):
        d = ref_aff.translate(datum, ty, tx)
        out[f"translate_{i}_mask"] = d["cine_lv_myo_masks_merged"]
        out[f"translate_{i}_strain"] = d["StrainInfo"]["CCmid"]
        out[f"translate_{i}_tos"] = d["TOSAnalysis"]["TOSfullRes_Jerry"]
    ns = [0, 1, 5, -3, 126, 130]
    out["rot_n"] = np.array(ns, dtype=np.int64)
    angles = []
    for i, n in enumerate(ns):
        SKROTATE_CALLS.clear()
        d = ref_aff.rotate(datum, n)
        assert len(SKROTATE_CALLS) == 1, SKROTATE_CALLS
        ang, kw = SKROTATE_CALLS[0]
        assert kw == {"resize": False, "preserve_range": True, "order": 0}, kw
        angles.append(ang)
        out[f"rotate_{i}_strain"] = d["StrainInfo"]["CCmid"]
        out[f"rotate_{i}_tos"] = d["TOSAnalysis"]["TOSfullRes_Jerry"]
    out["rot_angle_degree"] = np.array(angles, dtype=np.float64)
    out["skrotate_order"] = np.int64(0)
    np.savez_compressed(OUT_AUG, **out)
    print(f"wrote {OUT_AUG} ({OUT_AUG.stat().st_size} bytes, {len(out)} arrays)")


if __name__ == "__main__":
    main()

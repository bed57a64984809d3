"""Oracle for the callers either side of the hot path (TEST INFRASTRUCTURE).

Restates the in-tree [REF] pieces (pair split, frame alignment, the registration
loss) and wires the operator oracle into the ``forward_volume`` / pairwise
contracts the trainers consume.  ``tests/golden/make_golden.py`` checks the
[REF] restatements against the reference's own code imported from
/root/reference and commits the resulting vectors.
"""
from __future__ import annotations

import numpy as np
import torch

from .lddmm import Conventions, DEFAULT, FluidMetric, interp, shoot
from .strain import N_SECTORS, strain_matrix


def split_vol_to_registration_pairs(vol, split_method: str = "Lagrangian", output_dim: int = 3):
    """Pair construction, behaviour of /root/reference/modules/data/__init__.py:93-121 (golden-checked)."""
    n_slices, n_chan, n_t, hh, ww = vol.shape
    assert n_t > 1, f"n_frames should be larger than 1, but got {n_t}"
    t_idx = torch.arange(1, n_t)
    if split_method == "Lagrangian":
        s_idx = torch.zeros(n_t - 1, dtype=torch.long)
    elif split_method == "Eulerian":
        s_idx = t_idx - 1
    else:
        raise ValueError(f"Unrecognized split_method: {split_method}")
    src, tar = vol.index_select(2, s_idx), vol.index_select(2, t_idx)
    if output_dim == 2:
        src = src.reshape(-1, n_chan, hh, ww)
        tar = tar.reshape(-1, n_chan, hh, ww)
    return src, tar


def align_n_frames_to(volume: np.ndarray, n_target_frames: int, frame_idx: int = -1,
                      padding_method: str = "edge"):
    """Frame alignment, behaviour of /root/reference/modules/data/datareader/DENSE_IO_utils.py:2-46 (golden-checked)."""
    moved = np.moveaxis(volume, frame_idx, -1)
    n = moved.shape[-1]
    if n >= n_target_frames:
        out = moved[..., :n_target_frames]
    else:
        out = np.pad(moved, [(0, 0)] * (moved.ndim - 1) + [(0, n_target_frames - n)], mode=padding_method)
    return np.moveaxis(out, -1, frame_idx)


def forward_pairs(v0, src, tar, metric: FluidMetric, num_steps: int = 10,
                  conv: Conventions = DEFAULT):
    """Pairwise contract of /root/reference/modules/trainer/reg_trainer.py:45,222-225.

    v0: (P,2,H,W) initial velocity (stands in for the missing network), src/tar (P,1,H,W).
    """
    m0, vel, u = shoot(metric, v0, num_steps, conv=conv)
    return {
        "displacement": u,
        "velocity": vel,
        "momentum": m0,
        "deformed_source": interp(src, u, 1.0, conv),
    }


def forward_volume(v0, src_vol, tar_vol, metric: FluidMetric, num_steps: int = 10,
                   n_sectors: int = N_SECTORS, n_frames: int | None = 40,
                   conv: Conventions = DEFAULT, theta0=None, clockwise=None):
    """``forward_volume`` contract of joint_registration_strainmat_LMA.py:307,314-318.

    v0: (B*(T-1), 2, H, W) ordered slice-major; src_vol, tar_vol: (B,1,T-1,H,W).
    """
    B, C, T1, H, W = tar_vol.shape
    assert C == 1
    src = src_vol.reshape(B * T1, 1, H, W)
    out = forward_pairs(v0, src, tar_vol.reshape(B * T1, 1, H, W), metric, num_steps, conv)
    u = out["displacement"]
    S = strain_matrix(u.reshape(B, T1, 2, H, W), tar_vol[:, 0], src_vol[:, 0, 0],
                      n_sectors, n_frames, conv, theta0=theta0, clockwise=clockwise)
    return {
        "strain_matrix": S,
        "deformed_source": out["deformed_source"].reshape(B, 1, T1, H, W),
        "velocity": out["velocity"],
        "momentum": out["momentum"],
        "displacement": u,
    }


def registration_loss_terms(pred, tar):
    """Per-pair sums behind the loss below: (P,2) = {sum (tar - Sdef)^2, sum v.m} (float64 accumulation)."""
    Sdef = pred["deformed_source"]
    P = pred["velocity"].shape[0]
    sq = ((tar.reshape(P, -1).double() - Sdef.reshape(P, -1).double()) ** 2).sum(dim=1)
    vm = (pred["velocity"].double() * pred["momentum"].double()).reshape(P, -1).sum(dim=1)
    return torch.stack([sq, vm], dim=1)


def registration_reconstruction_loss(pred, target, sigma=0.03, regularization_weight=0.1):
    """/root/reference/modules/loss/registration_losses.py:22-28."""
    Sdef = pred["deformed_source"]
    tar = target["registration_target"]
    recon = torch.mean((tar - Sdef) ** 2)
    reg = (pred["velocity"] * pred["momentum"]).sum() / tar.numel()
    return 0.5 * recon / (sigma * sigma) + reg * regularization_weight


def merge_data_of_same_slice_from_batch(batch, reg_pred_dict, n_frames_to_use_for_regression):
    """Slice regrouping, behaviour of /root/reference/modules/trainer/joint_registration_regression_trainer.py:54-120
    (golden-checked): per slice, the displacement fields of its pairs in batch order -> (2, F, H, W), cropped or
    zero-padded to F frames; labels from the slice's first pair.  Slices in order of first appearance."""
    ids = list(batch["slice_full_id"])
    order = list(dict.fromkeys(ids))
    F = int(n_frames_to_use_for_regression)
    fields, first = [], []
    for sid in order:
        idx = [i for i, x in enumerate(ids) if x == sid]
        first.append(idx[0])
        d = torch.stack([reg_pred_dict["displacement"][i] for i in idx], dim=1)[:, :F]      # (2, n, H, W)
        if d.shape[1] < F:
            d = torch.cat([d, d.new_zeros(d.shape[0], F - d.shape[1], *d.shape[2:])], dim=1)
        fields.append(d)
    ft = torch.tensor(first)
    return {
        "pred_displacement_fields": torch.stack(fields, dim=0),
        "TOS": batch["TOS"][ft],
        "sector_LMA_labels": batch["sector_LMA_labels"][ft],
        "slice_LMA_label": torch.tensor([batch["slice_LMA_label"][i].item() for i in first]),
        "batch_slice_full_ids": order,
    }

"""Host-side callers either side of the hot path (mirrors of the reference's own helpers).

* ``split_vol_to_registration_pairs``  behaviour of /root/reference/modules/data/__init__.py:93-121
* ``align_n_frames_to``                behaviour of /root/reference/modules/data/datareader/DENSE_IO_utils.py:2-46

* ``merge_data_of_same_slice_from_batch``  behaviour of
  /root/reference/modules/trainer/joint_registration_regression_trainer.py:54-120, with the displacement
  regrouping done by one device kernel (``b2_regroup_pairs``) instead of per-slice stack / permute / pad.

Same names, arguments and error behaviour, written independently; all are checked against the reference's own
code through the committed golden vectors (tests/golden/ref_boundary.npz, ref_regroup.npz).
"""
from __future__ import annotations

import numpy as np
import torch

_SPLITS = ("Lagrangian", "Eulerian")


def split_vol_to_registration_pairs(vol, split_method: str = "Lagrangian", output_dim: int = 3):
    """Cut a cine volume (B,C,T,H,W) into T-1 (source, target) frame pairs per slice.

    Lagrangian: every target frame 1..T-1 is registered to frame 0; Eulerian: frame t to frame t+1.
    ``output_dim=3`` keeps the (B,C,T-1,H,W) layout, ``output_dim=2`` folds pairs into the batch axis.
    The Lagrangian source is a stride-0 view of frame 0 (values identical to the reference's ``repeat``): the
    fused kernel indexes frame 0 per slice, so the T-1 copies are never materialised.
    """
    B, C, T, H, W = vol.shape
    assert T > 1, f"n_frames should be larger than 1, but got {T}"
    if split_method not in _SPLITS:
        raise ValueError(f"Unrecognized split_method: {split_method}")
    targets = vol.narrow(2, 1, T - 1)
    if split_method == "Lagrangian":
        sources = vol.narrow(2, 0, 1).expand(B, C, T - 1, H, W)
    else:
        sources = vol.narrow(2, 0, T - 1)
    if output_dim == 2:
        fold = (B * (T - 1), C, H, W)
        return sources.reshape(fold), targets.reshape(fold)
    return sources, targets


def align_n_frames_to(volume, n_target_frames, frame_idx=-1, padding_method="edge"):
    """Bring the frame axis of a numpy volume to exactly ``n_target_frames``.

    Longer volumes keep their first frames; shorter ones are padded at the end with ``numpy.pad`` in
    ``padding_method`` mode ('edge' repeats the last frame), so an unknown mode raises numpy's ``ValueError``.
    """
    axis = frame_idx % volume.ndim
    have = volume.shape[axis]
    if have >= n_target_frames:
        return volume[(slice(None),) * axis + (slice(0, n_target_frames),)]
    widths = [(0, n_target_frames - have) if ax == axis else (0, 0) for ax in range(volume.ndim)]
    return np.pad(volume, widths, mode=padding_method)


class _RegroupPairs(torch.autograd.Function):
    """(P,C,H,W) per-pair fields -> (n_slices,C,F,H,W) by slot table; differentiable like the reference's
    stack / permute / pad (the joint trainer backpropagates the LMA loss through it,
    /root/reference/modules/trainer/joint_registration_regression_trainer.py:290-320)."""

    @staticmethod
    def forward(ctx, u, slot_d, n_slices, F):
        from . import _lib
        from ._lib import check, lib, ptr, stream
        P, C, H, W = u.shape
        with _lib.on_device(u):
            out = torch.empty((n_slices, C, F, H, W), dtype=u.dtype, device=u.device)
            check(lib().b2_regroup_pairs(ptr(u), ptr(slot_d), ptr(out), P, n_slices, F, C, H, W, stream()),
                  "b2_regroup_pairs")
        _lib.count_launch()
        ctx.save_for_backward(slot_d)
        ctx.dims = (P, C, H, W, n_slices, F)
        return out

    @staticmethod
    def backward(ctx, gout):
        from . import _lib
        from ._lib import check, lib, ptr, stream
        (slot_d,) = ctx.saved_tensors
        P, C, H, W, n_slices, F = ctx.dims
        gout = gout.contiguous()
        with _lib.on_device(gout):
            gu = torch.empty((P, C, H, W), dtype=gout.dtype, device=gout.device)
            check(lib().b2_regroup_pairs_bwd(ptr(gout), ptr(slot_d), ptr(gu), P, n_slices, F, C, H, W, stream()),
                  "b2_regroup_pairs_bwd")
        _lib.count_launch()
        return gu, None, None, None


def merge_data_of_same_slice_from_batch(batch, reg_pred_dict, n_frames_to_use_for_regression, used_device):
    """Regroup the per-pair outputs of a batch by slice.

    ``batch['slice_full_id']`` names the slice of every pair; the displacement fields (P,2,H,W) of a slice are laid
    out in batch order as frames of a (2, F, H, W) block, cropped to ``F = n_frames_to_use_for_regression`` or
    zero-padded at the end; ``TOS`` / ``sector_LMA_labels`` / ``slice_LMA_label`` are taken from the first pair of
    each slice.  Returns the reference's dict.  Slices are ordered by first appearance (the reference iterates a
    ``set``, i.e. in arbitrary order; ``'batch_slice_full_ids'`` names the order used).
    """
    from . import _lib
    from ._lib import require_cuda
    ids = list(batch["slice_full_id"])
    u = reg_pred_dict["displacement"].contiguous()
    require_cuda(u)
    P, C, H, W = u.shape
    if len(ids) != P:
        raise _lib.B2Error(f"{len(ids)} slice ids for {P} pairs")
    F = int(n_frames_to_use_for_regression)
    order, seen, first, slot = [], {}, [], []
    for p, sid in enumerate(ids):
        if sid not in seen:
            seen[sid] = [len(order), 0]
            order.append(sid)
            first.append(p)
        s, pos = seen[sid]
        slot.append(s * F + pos if pos < F else -1)
        seen[sid][1] = pos + 1
    slot_d = torch.tensor(slot, dtype=torch.int32).to(u.device)
    out = _RegroupPairs.apply(u, slot_d, len(order), F)      # keeps the graph: the LMA loss reaches the registration net
    first_t = torch.tensor(first)
    return {
        "pred_displacement_fields": out,
        "TOS": batch["TOS"][first_t].to(used_device),
        "sector_LMA_labels": batch["sector_LMA_labels"][first_t].to(used_device),
        "slice_LMA_label": torch.tensor([batch["slice_LMA_label"][i].item() for i in first]).to(used_device),
        "batch_slice_full_ids": order,
    }

"""Host-side callers either side of the hot path (mirrors of the reference's own helpers).

* ``split_vol_to_registration_pairs``  behaviour of /root/reference/modules/data/__init__.py:93-121
* ``align_n_frames_to``                behaviour of /root/reference/modules/data/datareader/DENSE_IO_utils.py:2-46

Same names, arguments and error behaviour, written independently; both are checked against the reference's own
code through the committed golden vectors (tests/golden/ref_boundary.npz).
"""
from __future__ import annotations

import numpy as np

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

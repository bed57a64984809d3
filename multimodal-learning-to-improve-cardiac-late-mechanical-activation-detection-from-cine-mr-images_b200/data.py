"""Host-side callers either side of the hot path (mirrors of the reference's own helpers).

* ``split_vol_to_registration_pairs``  /root/reference/modules/data/__init__.py:93-121
* ``align_n_frames_to``                /root/reference/modules/data/datareader/DENSE_IO_utils.py:2-46

Same names, arguments and error behaviour; checked against the reference's own
code through the committed golden vectors (tests/golden/ref_boundary.npz).
"""
from __future__ import annotations

import numpy as np


def split_vol_to_registration_pairs(vol, split_method: str = "Lagrangian", output_dim: int = 3):
    """vol (B,C,T,H,W) -> (src, tar), each (B,C,T-1,H,W) or flattened to (B*(T-1),C,H,W)."""
    batch_size, n_channels, n_frames, height, width = vol.shape
    assert n_frames > 1, f"n_frames should be larger than 1, but got {n_frames}"
    if split_method == "Lagrangian":
        # expand (a view) instead of repeat: the fused kernel indexes frame 0 per slice, so the
        # T-1 copies are never materialised; values are identical to the reference's repeat.
        src = vol[:, :, :1].expand(batch_size, n_channels, n_frames - 1, height, width)
        tar = vol[:, :, 1:]
    elif split_method == "Eulerian":
        src = vol[:, :, :-1]
        tar = vol[:, :, 1:]
    else:
        raise ValueError(f"Unrecognized split_method: {split_method}")
    if output_dim == 2:
        src = src.reshape(batch_size * (n_frames - 1), n_channels, height, width)
        tar = tar.reshape(batch_size * (n_frames - 1), n_channels, height, width)
    return src, tar


def align_n_frames_to(volume, n_target_frames, frame_idx=-1, padding_method="edge"):
    """Crop to the first ``n_target_frames`` frames or pad at the end (numpy semantics)."""
    n_frames = volume.shape[frame_idx]
    if n_frames >= n_target_frames:
        indices = [slice(None)] * volume.ndim
        indices[frame_idx] = slice(0, n_target_frames)
        return volume[tuple(indices)]
    paddings = [(0, 0)] * volume.ndim
    paddings[frame_idx] = (0, n_target_frames - n_frames)
    return np.pad(volume, paddings, mode=padding_method)

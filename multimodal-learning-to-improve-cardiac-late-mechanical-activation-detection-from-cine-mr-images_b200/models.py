"""``models``-compatible shim: the modules the reference's trainers instantiate.

The reference does ``from models import build_model`` (/root/reference/main.py:42-46)
but ships no ``models`` package, so the networks that
*produce* the initial velocity are stand-ins (plain torch/cuDNN - library code,
outside the measured hot path).  What is pinned by the reference and kept
exactly is the interface: ``build_model(model_config)`` dispatching on
``model_config['type']`` (/root/reference/configs/config.json:110,117),
``forward_volume(src_vol, tar_vol)`` and its dict keys
(joint_registration_strainmat_LMA.py:307,314-318), the pairwise
``model(src, tar)`` dict and ``.sigma`` (reg_trainer.py:45,222-225,230) and
``LMA_model(strain_matrix) -> {'TOS': (B,126)}`` (joint_registration_strainmat_LMA.py:308).
Everything between ``v0`` and the returned dict is the B200-native hot path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .ops import FluidMetric
from .shooting import shoot_warp_pairs, shoot_warp_strain


class VelocityNet(nn.Module):
    """Small encoder-decoder mapping a (src, tar) pair to an initial velocity v0 (P,2,H,W).

    Stand-in for the network the reference does not ship.  All feature maps live at 1/2 and 1/4 resolution and the
    2-channel velocity is upsampled at the end: a batch of 1536 pairs at 128x128 makes every full-resolution
    16-channel activation 1.6 GB, so a full-resolution design is HBM-bound on activations alone (~100 ms per
    training step on a B200 against 13 ms for the whole registration path).
    """

    def __init__(self, width: int = 16, max_velocity: float = 3.0):
        super().__init__()
        self.enc1 = nn.Conv2d(2, width, 3, stride=2, padding=1)             # 1/2
        self.enc2 = nn.Conv2d(width, 2 * width, 3, stride=2, padding=1)     # 1/4
        self.mid = nn.Conv2d(2 * width, 2 * width, 3, padding=1)
        self.dec1 = nn.Conv2d(3 * width, width, 3, padding=1)               # 1/2, skip from enc1
        self.out = nn.Conv2d(width, 2, 3, padding=1)
        self.max_velocity = max_velocity
        nn.init.normal_(self.out.weight, std=1e-3)
        nn.init.zeros_(self.out.bias)

    @staticmethod
    def _up2(x, size):
        """Nearest-neighbour x2 as one expand + copy (ATen's bilinear upsample kernel loops over batch*channels
        inside each thread: 12 ms per call at 1536 pairs); the 3x3 conv / the fluid metric that follow smooth it."""
        n, c, h, w = x.shape
        y = x[:, :, :, None, :, None].expand(n, c, h, 2, w, 2).reshape(n, c, 2 * h, 2 * w)
        return y[..., : size[0], : size[1]]

    def forward(self, src, tar):
        x = torch.cat([src, tar], dim=1).contiguous(memory_format=torch.channels_last)   # cuDNN's native layout
        e1 = F.relu(self.enc1(x))
        e2 = F.relu(self.enc2(e1))
        m = F.relu(self.mid(e2))
        up = self._up2(m, e1.shape[-2:]).contiguous(memory_format=torch.channels_last)
        d1 = F.relu(self.dec1(torch.cat([up, e1], dim=1)))
        v = self.max_velocity * torch.tanh(self.out(d1))
        return self._up2(v, x.shape[-2:]).contiguous()


def svd_smooth(S: torch.Tensor, rank: int, grad: str = "projection") -> torch.Tensor:
    """Rank truncation of each (sectors x frames) strain matrix.

    Same operation as ``SVDDenoise`` (/root/reference/modules/data/utils/DENSE_utils.py:11-14: ``u @ diag(s[:rank]) @
    vh``), selected by ``strainmat_smoothing_method: "SVD"`` (configs/config.json:113-114); checked against the
    reference function's own output (tests/golden/ref_sectors.npz).

    The reference applies it to numpy ground truth, so it defines no gradient.  Two explicit choices here:

    * ``grad="projection"`` (default, intended): the left basis ``U_r`` is treated as a constant, the backward is
      ``g -> U_r U_r^T g``, the orthogonal projection onto the retained subspace.  It is defined for every input -
      the strain matrix is edge-padded beyond the cine frames, hence rank deficient with repeated zero singular
      values, where the derivative of the SVD itself does not exist (NaN in ``torch.linalg.svd`` backward).
    * ``grad="exact"``: differentiate through ``torch.linalg.svd`` (the true derivative of the truncation; needs
      distinct singular values).
    """
    if grad == "exact":
        U, s, Vh = torch.linalg.svd(S, full_matrices=False)
        return (U[..., :rank] * s[..., None, :rank]) @ Vh[..., :rank, :]
    if grad != "projection":
        raise ValueError(f"grad must be 'projection' or 'exact', got {grad!r}")
    # U_r U_r^T S equals the rank-r truncation U_r diag(s_r) V_r^T
    with torch.no_grad():
        U = torch.linalg.svd(S, full_matrices=False)[0][..., :rank]
    return U @ (U.transpose(-1, -2) @ S)


class JointRegisterStrainMatNet(nn.Module):
    """Registration network + geodesic shooting + strain matrix (``forward_volume``)."""

    def __init__(self, config: dict | None = None):
        super().__init__()
        config = dict(config or {})
        self.n_strain_matrix_frames = int(config.get("n_strain_matrix_frames", 40))
        self.n_sectors = int(config.get("n_sectors", 126))
        self.num_steps = int(config.get("num_steps", 10))
        self.sigma = float(config.get("sigma", 0.03))
        self.smoothing = config.get("strainmat_smoothing_method", None)
        self.smoothing_rank = int(config.get("strainmat_smoothing_SVD_rank", 5))
        self.smoothing_grad = config.get("strainmat_smoothing_grad", "projection")       # see svd_smooth
        self.fused_loss_terms = bool(config.get("fused_loss_terms", False))   # adds 'registration_loss_terms'
        self.metric = FluidMetric(config.get("fluid_params", (1.0, 0.1, 0.05)))
        self.velocity_net = VelocityNet(int(config.get("velocity_net_width", 16)),
                                        float(config.get("max_velocity", 3.0)))

    def forward(self, src, tar):
        """Pairwise contract: src, tar (P,1,H,W) -> displacement / velocity / momentum / deformed_source."""
        v0 = self.velocity_net(src, tar).float()      # the path is fp32 even when the net runs under autocast
        return shoot_warp_pairs(v0, src, tar, self.metric, self.num_steps, loss_terms=self.fused_loss_terms)

    def forward_volume(self, src_vol, tar_vol, theta0=None, clockwise=None):
        """src_vol, tar_vol (B,1,T-1,H,W) -> {'strain_matrix','deformed_source','velocity','momentum',...}.
        ``theta0`` / ``clockwise``: optional per-slice sector frame of the strain-matrix rows (strain.py)."""
        B, C, T1, H, W = tar_vol.shape
        # pair (b, t) registers src_vol[b, :, t] to tar_vol[b, :, t]: frame 0 for every t under the Lagrangian
        # split, frame t under the Eulerian one (modules/data/__init__.py:108-113)
        v0 = self.velocity_net(src_vol.reshape(B * T1, C, H, W), tar_vol.reshape(B * T1, C, H, W)).float()   # path is fp32
        out = shoot_warp_strain(v0, src_vol, tar_vol, self.metric, self.num_steps,
                                n_sectors=self.n_sectors, n_frames=self.n_strain_matrix_frames,
                                loss_terms=self.fused_loss_terms, theta0=theta0, clockwise=clockwise)
        if self.smoothing == "SVD":
            out["strain_matrix"] = svd_smooth(out["strain_matrix"], self.smoothing_rank, self.smoothing_grad)
        return out


class NetStrainMat2LMA(nn.Module):
    """Strain matrix (B,1,126,40) -> {'TOS': (B,126)}  (configs/config.json:116-125)."""

    def __init__(self, config: dict | None = None):
        super().__init__()
        config = dict(config or {})
        n_layers = int(config.get("num_conv_layers", 3))
        ch = int(config.get("inner_conv_channel_num", 16))
        cin = int(config.get("input_channel_num", 1))
        self.n_frames = int(config.get("n_frames", 40))
        self.n_sectors = int(config.get("n_sectors", 126))
        self.n_classes = int(config.get("n_classes", 1))
        layers = []
        for i in range(n_layers):
            layers += [nn.Conv2d(cin if i == 0 else ch, ch, 3, padding=1), nn.ReLU(inplace=True)]
        self.features = nn.Sequential(*layers)
        self.head = nn.Conv2d(ch, self.n_classes, (1, self.n_frames))

    def forward(self, strain_matrix):
        x = self.features(strain_matrix)
        y = self.head(x)                        # (B, n_classes, n_sectors, 1)
        tos = y.squeeze(-1)
        if self.n_classes == 1:
            tos = tos.squeeze(1)                # (B, n_sectors)
        return {"TOS": tos}


_REGISTRY = {
    "JointRegisterStrainMatNet": JointRegisterStrainMatNet,
    "NetStrainMat2LMA": NetStrainMat2LMA,
}


def build_model(model_config: dict) -> nn.Module:
    """``models.build_model`` (/root/reference/main.py:42-46): dispatch on ``model_config['type']``."""
    mtype = model_config["type"]
    if mtype not in _REGISTRY:
        raise NotImplementedError(f"model type {mtype!r} not implemented")
    return _REGISTRY[mtype](model_config)

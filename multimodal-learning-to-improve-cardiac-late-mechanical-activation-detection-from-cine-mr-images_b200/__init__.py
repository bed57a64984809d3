"""B200-native registration-to-strain hot path (lagomorph-style LDDMM shooting + strain).

Drop-in for the operator surface the reference's trainers use
(``lagomorph.interp / splat / FluidMetric / expmap / EPDiff_step / Ad_star /
jacobian_times_vectorfield / compose_disp_vel`` and the ``models`` package's
``build_model / forward_volume``); the arithmetic runs in hand-written sm_100a
CUDA behind the C ABI of ``include/b2lddmm.h``.  No CPU fallback.
"""
from . import _lib, augment, data, losses, models, ops, parallel, shooting, strain, synthetic  # noqa: F401
from .losses import RegistrationReconstructionLoss, reconstruction_loss_from_terms  # noqa: F401
from .models import JointRegisterStrainMatNet, NetStrainMat2LMA, build_model  # noqa: F401
from .ops import (  # noqa: F401
    Ad_star,
    FluidMetric,
    compose_disp_vel,
    interp,
    jacobian_times_vectorfield,
    splat,
)
from .shooting import EPDiff_step, HostPipeline, expmap, shoot_warp_pairs, shoot_warp_strain  # noqa: F401
from .strain import sector_map, strain_matrix  # noqa: F401

__version__ = "0.1.0"

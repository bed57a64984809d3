"""Drop-in for ``import lagomorph as lm`` (the reference imports it at
/root/reference/modules/trainer/joint_registration_strainmat_LMA.py:5, reg_trainer.py:4, ...).

Put ``<repo>/drop_in`` on PYTHONPATH; every name resolves to the B200 CUDA implementation.
"""
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from _load import load as _load  # noqa: E402

_pkg = _load()

interp = _pkg.interp
splat = _pkg.splat
FluidMetric = _pkg.FluidMetric
expmap = _pkg.expmap
EPDiff_step = _pkg.EPDiff_step
Ad_star = _pkg.Ad_star
jacobian_times_vectorfield = _pkg.jacobian_times_vectorfield
compose_disp_vel = _pkg.compose_disp_vel

__all__ = ["interp", "splat", "FluidMetric", "expmap", "EPDiff_step", "Ad_star",
           "jacobian_times_vectorfield", "compose_disp_vel"]

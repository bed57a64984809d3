"""Drop-in for the reference's missing ``models`` package (``from models import build_model``,
/root/reference/main.py:42-46).  Put ``<repo>/drop_in`` on PYTHONPATH."""
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from _load import load as _load  # noqa: E402

_pkg = _load()

build_model = _pkg.build_model
JointRegisterStrainMatNet = _pkg.JointRegisterStrainMatNet
NetStrainMat2LMA = _pkg.NetStrainMat2LMA

__all__ = ["build_model", "JointRegisterStrainMatNet", "NetStrainMat2LMA"]

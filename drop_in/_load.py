"""Locate and import the B200 package (hyphenated directory name) for the drop-in shims."""
import importlib
import pathlib
import sys

_ROOT = pathlib.Path(__file__).resolve().parent.parent
_PKG = ("multimodal-learning-to-improve-cardiac-late-mechanical-activation-"
        "detection-from-cine-mr-images_b200")


def load():
    if str(_ROOT) not in sys.path:
        sys.path.insert(0, str(_ROOT))
    pkg = importlib.import_module(_PKG)
    sys.modules.setdefault("b2lddmm", pkg)
    return pkg

"""idee_b200 -- B200-native (sm_100a) implementation of the IDEE hot path.

Swin-3D encoder -> LFQ binary-driver quantiser -> CNN-3D extremes classifier (+ the training-step losses), behind the
reference's own constructors / forward contracts.  The compute lives in ``idee_b200/lib/libidee_b200.so`` (hand-written
CUDA, C ABI in ``include/idee_b200.h``); Python is plumbing.  There is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import sys

__version__ = "0.1.0"


def install_as_reference_modules() -> None:
    """Make ``importlib.import_module('models.encoder.Swin_3D')`` etc. resolve to the idee_b200 modules, so the
    reference's ``models/build.py::import_class`` (build.py:17-20) and training scripts pick up the CUDA path."""
    import importlib
    for name in ("encoder.Swin_3D", "encoder.CNN_3D", "codebook.LFQ", "classifier.CNN_3D", "losses", "build"):
        sys.modules["models." + name] = importlib.import_module("idee_b200.models." + name)

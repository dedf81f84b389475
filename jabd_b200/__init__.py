"""Import alias for the product package.

The product package directory is named after the reference repository
(``jabd-joint-attention-based-detector-for-small-face-detection_b200``), which
is not a valid Python identifier.  ``import jabd_b200`` resolves every
submodule from that directory, so ``jabd_b200.batched`` *is*
``<that dir>/batched.py`` -- there is no second copy of the code.
"""
import os as _os

_REAL = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "jabd-joint-attention-based-detector-for-small-face-detection_b200",
)
if not _os.path.isdir(_REAL):  # pragma: no cover - broken checkout
    raise ImportError("product package directory missing: " + _REAL)
__path__ = [_REAL]

PACKAGE_DIR = _REAL
REPO_ROOT = _os.path.dirname(_REAL)
__version__ = "0.1.0"

"""swinfuse: B200-native kernels for the Swin-UNet fusion hot path (host side).

``swinfuse.ops`` wraps the C ABI of ``libswinfuse.so``; the reference-compatible module files
live in ``../dropin`` (same file and class names as the reference: a001_WindowAttention.py ...
a013_ModelDefinition.py) and are imported by putting that directory on ``sys.path``:

    import swinfuse; swinfuse.install_dropin()      # then: from a013_ModelDefinition import MyModel
"""
import os
import sys

from . import ops  # noqa: F401
from ._lib import LIB_PATH, SwinFuseError, load  # noqa: F401
from .ops import get_default_precision, set_default_precision  # noqa: F401

DROPIN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin")

_DROPIN_MODULES = ("a001_WindowAttention", "a002_AutoPathWinAtt", "a003_AutoPathMLP",
                   "a004_AddAndLayerNormWithOtherModule", "a005_BasicBlock", "a006_PaddingOperation", "a007_utils", "a008_loss",
                   "a009_NormalAndShiftWinsBlockPair", "a010_StateRecorder", "a011_PatchOperation",
                   "a012_SelfAndCrossBlockPair", "a013_ModelDefinition")


def install_dropin() -> str:
    """Put the drop-in module directory first on sys.path (ahead of the reference checkout) and
    forget any already-imported reference modules of the same names."""
    if DROPIN_DIR in sys.path:
        sys.path.remove(DROPIN_DIR)
    sys.path.insert(0, DROPIN_DIR)
    for m in _DROPIN_MODULES:
        mod = sys.modules.get(m)
        if mod is not None and os.path.dirname(os.path.abspath(getattr(mod, "__file__", ""))) != DROPIN_DIR:
            del sys.modules[m]
    return DROPIN_DIR


if os.environ.get("SWINFUSE_PRECISION"):
    set_default_precision(os.environ["SWINFUSE_PRECISION"])

# the tcgen05 bf16 path passes the 2e-2 parity tests (tests/test_gpu_bf16.py): bench.py uses it
BF16_READY = True

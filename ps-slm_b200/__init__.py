"""ps-slm_b200 — B200-native (sm_100a) implementation of the TASU speech→LLM bridge hot path
of PigeonDan1/ps-slm: CTC posterior → collapse compression → projector → splice.

The directory name carries the reference's hyphen, so it is imported through the
``ps_slm_b200`` shim at the repository root (``import ps_slm_b200``).
"""
from . import _lib
from ._lib import TasuError, build, lib

__all__ = ["_lib", "TasuError", "build", "lib"]

"""video-stab_b200: B200-native drop-in for the reference's per-frame stabilization hot path."""
from ._capi import LIB_PATH, VsError, lib  # noqa: F401  (raises ImportError if the CUDA library is missing)
from .stabilizer import Parameters, Stabilizer, StabilizerBatch  # noqa: F401
from . import kernels  # noqa: F401
from . import offline  # noqa: F401
from .roll import RollCorrection, RollParameters  # noqa: F401
from .autozoom import AutoZoomCrop  # noqa: F401
from .canvas import VirtualCanvas  # noqa: F401

__all__ = ["lib", "LIB_PATH", "VsError", "Parameters", "Stabilizer", "StabilizerBatch", "kernels", "offline", "RollCorrection", "RollParameters", "AutoZoomCrop", "VirtualCanvas"]

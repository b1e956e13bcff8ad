"""Driver for oracle/_ref/libstages_ref.so — the reference's own RollCorrection.cpp and AutoZoomCrop.cpp (SURVEY.md section 8f ranks
1, 2), compiled unmodified by oracle/build_ref.py.  TEST INFRASTRUCTURE ONLY.

Both components are GPU-only in the reference (cv::cuda::*); here every cv::cuda:: call is served by the CPU function of the
same OpenCV (cv2 wheel) — see oracle/mini_cv/opencv2/mini_cv_cuda.hpp for the mapping and the stated residuals.  RollCorrection
keeps its state in file-scope statics (RollCorrection.cpp:13-14), so every `RefRollCorrection` loads a private copy of the library.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os

import numpy as np

from . import ref_lib

LIB = os.path.join(ref_lib.HERE, "_ref", "libstages_ref.so")


def available() -> bool:
    try:
        from . import build_ref
        build_ref.build()
    except Exception:
        pass
    return os.path.exists(LIB)


@dataclasses.dataclass
class RollParameters:
    """vs::RollCorrection::Parameters, include/video/RollCorrection.h:16-38"""
    scaleFactor: float = 0.25
    cannyThresholdLow: float = 50.0
    cannyThresholdHigh: float = 150.0
    cannyAperture: int = 3
    houghRho: float = 1.0
    houghTheta: float = float(np.float32(np.pi / 180.0))
    houghThreshold: int = 100
    angleFilterMin: float = -10.0
    angleFilterMax: float = 10.0
    angleSmoothingAlpha: float = 0.1
    angleDecay: float = 0.995
    maxAngleChangeDeg: float = 0.5


def _declare(lib):
    vp, ci, cz, PI = C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_int)
    sig = {
        "mini_cv_set_ops": (None, [C.POINTER(ref_lib.OpsTable)]),
        "vsroll_params_new": (vp, []), "vsroll_params_delete": (None, [vp]),
        "vsroll_params_set": (ci, [vp, C.c_char_p, C.c_double]),
        "vsroll_params_get": (ci, [vp, C.c_char_p, C.POINTER(C.c_double)]),
        "vsroll_correct": (ci, [vp, vp, ci, ci, cz, vp, cz, PI, PI]),
        "vsroll_smoothed_angle": (C.c_double, []), "vsroll_first_frame": (ci, []),
        "vszoom_crop": (ci, [vp, ci, ci, cz, C.c_double, vp, cz, PI, PI]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


class _Stage:
    def __init__(self, use_optimized: bool = False):
        self.lib = _declare(ref_lib.load_private_copy(LIB))
        self.ops = ref_lib.CvOps(use_optimized)
        self.lib.mini_cv_set_ops(C.byref(self.ops.table))
        self.trace: list = []
        self.ops.log = lambda op, kw: self.trace.append((op, kw))
        self._out = None

    def _buf(self, n):
        if self._out is None or self._out.size < n:
            self._out = np.empty(n, np.uint8)
        return self._out


class RefRollCorrection(_Stage):
    """vs::RollCorrection::autoCorrectRoll with its own (fresh) static state."""

    def __init__(self, params: RollParameters | None = None, use_optimized: bool = False):
        super().__init__(use_optimized)
        self.params = params or RollParameters()
        self._p = self.lib.vsroll_params_new()
        for k, v in vars(self.params).items():
            if self.lib.vsroll_params_set(self._p, k.encode(), float(v)) != 0:
                raise KeyError(k)

    def correct(self, frame: np.ndarray):
        frame = np.ascontiguousarray(frame)
        h, w = frame.shape[:2]
        out = self._buf(h * w * 3)
        ow, oh = C.c_int(), C.c_int()
        self.trace.clear()
        rc = self.lib.vsroll_correct(self._p, frame.ctypes.data, w, h, frame.strides[0], out.ctypes.data, out.size, C.byref(ow), C.byref(oh))
        if rc < 0:
            raise RuntimeError(f"reference autoCorrectRoll failed rc={rc}: {self.ops.errors[-3:]}")
        if rc == 0:
            return None
        return out[: ow.value * oh.value * 3].reshape(oh.value, ow.value, 3).copy()

    @property
    def smoothed_angle(self) -> float:
        return float(self.lib.vsroll_smoothed_angle())

    def last(self, op: str):
        for name, kw in reversed(self.trace):
            if name == op:
                return kw
        return None


class RefAutoZoomCrop(_Stage):
    def crop(self, frame: np.ndarray, margin: float = 0.05):
        frame = np.ascontiguousarray(frame)
        h, w = frame.shape[:2]
        out = self._buf(max(h * w * 3, 640 * 360 * 3))
        ow, oh = C.c_int(), C.c_int()
        rc = self.lib.vszoom_crop(frame.ctypes.data, w, h, frame.strides[0], margin, out.ctypes.data, out.size, C.byref(ow), C.byref(oh))
        if rc < 0:
            raise RuntimeError(f"reference autoZoomCrop failed rc={rc}: {self.ops.errors[-3:]}")
        if rc == 0:
            return None
        return out[: ow.value * oh.value * 3].reshape(oh.value, ow.value, 3).copy()

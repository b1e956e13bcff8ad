"""vs::RollCorrection over the C-ABI (include/vstab_b200.h, vs_roll_*): the mirror of the reference's
`RollCorrection::autoCorrectRoll(input, params)` (include/video/RollCorrection.h:16-51).  The reference keeps its smoothed angle
in process-global statics; here the state belongs to the `RollCorrection` object."""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

from ._capi import VsRollParams, check, lib

_FIELDS = [("scaleFactor", "scale_factor"), ("cannyThresholdLow", "canny_threshold_low"), ("cannyThresholdHigh", "canny_threshold_high"),
           ("cannyAperture", "canny_aperture"), ("houghRho", "hough_rho"), ("houghTheta", "hough_theta"), ("houghThreshold", "hough_threshold"),
           ("angleFilterMin", "angle_filter_min"), ("angleFilterMax", "angle_filter_max"), ("angleSmoothingAlpha", "angle_smoothing_alpha"),
           ("angleDecay", "angle_decay"), ("maxAngleChangeDeg", "max_angle_change_deg")]


@dataclasses.dataclass
class RollParameters:
    """vs::RollCorrection::Parameters (RollCorrection.h:16-38)"""
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

    def to_c(self) -> VsRollParams:
        p = VsRollParams()
        for a, b in _FIELDS:
            setattr(p, b, getattr(self, a))
        return p

    @classmethod
    def from_c(cls, p: VsRollParams) -> "RollParameters":
        return cls(**{a: getattr(p, b) for a, b in _FIELDS})

    @classmethod
    def from_yaml(cls, path: str) -> "RollParameters":
        """the `roll_correction:` section of a reference config.yaml (keys of examples/vsg.cpp:988-1000)"""
        p = VsRollParams()
        check(lib.vs_roll_params_default(C.byref(p)))
        check(lib.vs_roll_params_from_yaml(path.encode(), C.byref(p)))
        return cls.from_c(p)

    @classmethod
    def from_yaml_string(cls, text: str) -> "RollParameters":
        p = VsRollParams()
        check(lib.vs_roll_params_default(C.byref(p)))
        check(lib.vs_roll_params_from_yaml_string(text.encode(), C.byref(p)))
        return cls.from_c(p)


class RollCorrection:
    def __init__(self, params: RollParameters | None = None, device: int = 0):
        self.params = params or RollParameters()
        self._h = C.c_void_p()
        cp = self.params.to_c()
        check(lib.vs_roll_create(C.byref(cp), device, C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.vs_roll_destroy(h)
            self._h = None

    def autoCorrectRoll(self, frame: np.ndarray):
        """HxWx3 uint8 BGR in, roll-corrected frame of the same size out (None for an empty frame)."""
        if frame is None or frame.size == 0:
            return None
        if frame.dtype != np.uint8 or frame.ndim != 3 or frame.shape[2] != 3:
            raise ValueError("frame must be HxWx3 uint8 (CV_8UC3 BGR)")
        frame = np.ascontiguousarray(frame)
        h, w = frame.shape[:2]
        out = np.empty_like(frame)
        check(lib.vs_roll_correct(self._h, frame.ctypes.data, w, h, frame.strides[0], out.ctypes.data, out.strides[0]))
        return out

    def correct_device(self, d_src: int, w: int, h: int, stride: int, d_dst: int, dst_stride: int, stream: int = 0):
        check(lib.vs_roll_correct_device(self._h, d_src, w, h, stride, d_dst, dst_stride, C.c_void_p(stream)))

    def reset(self):
        check(lib.vs_roll_reset(self._h))

    def state(self) -> dict:
        a, nl, ne, ln = C.c_double(), C.c_int(), C.c_int(), C.c_uint64()
        check(lib.vs_roll_state(self._h, C.byref(a), C.byref(nl), C.byref(ne), C.byref(ln)))
        return {"angle": a.value, "n_lines": nl.value, "n_edges": ne.value, "launches": ln.value}

    def debug(self) -> dict:
        """analysis image of the last frame: gray, edges, lines (rho, theta, votes) in cv::HoughLines order"""
        sw, sh = C.c_int(), C.c_int()
        check(lib.vs_roll_debug(self._h, C.byref(sw), C.byref(sh), None, None, None, 0))
        gray = np.empty((sh.value, sw.value), np.uint8)
        edges = np.empty((sh.value, sw.value), np.uint8)
        lines = np.zeros((4096, 3), np.float32)
        check(lib.vs_roll_debug(self._h, C.byref(sw), C.byref(sh), gray.ctypes.data, edges.ctypes.data, lines.ctypes.data, 4096))
        n = min(self.state()["n_lines"], 4096)
        return {"gray": gray, "edges": edges, "lines": lines[:n, :2].copy(), "votes": lines[:n, 2].astype(np.int32)}

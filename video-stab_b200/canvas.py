"""The virtual-canvas output stage on its own (Stabilizer::applyVirtualCanvasStabilization, reference
src/Stabilizer.cpp:2066-2443) over the C-ABI `vs_canvas_*`.  Inside `Stabilizer` the stage is switched on by
`Parameters(enableVirtualCanvas=True)`; this handle takes the frames and their corrections from the caller."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._capi import check, lib
from .stabilizer import Parameters


class VirtualCanvas:
    def __init__(self, params: Parameters | None = None, device: int = 0):
        self.params = params or Parameters(enableVirtualCanvas=True)
        self._h = C.c_void_p()
        cp = self.params.to_c()
        check(lib.vs_canvas_create(C.byref(cp), device, C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.vs_canvas_destroy(h)
            self._h = None

    def apply_device(self, d_src: int, w: int, h: int, stride: int, transform, d_dst: int, dst_stride: int, recent=None,
                     stream: int = 0):
        """One frame (device pointer) with its correction (dx, dy, da) -> the stage's output frame (device pointer, same
        size).  `recent`: the frame-to-frame transforms so far (n x 3), read by the first call only."""
        t = np.ascontiguousarray(np.asarray(transform, np.float32).reshape(3))
        r = np.zeros((0, 3), np.float32) if recent is None else np.ascontiguousarray(np.asarray(recent, np.float32).reshape(-1, 3))
        check(lib.vs_canvas_apply_device(self._h, d_src, w, h, stride, t.ctypes.data_as(C.POINTER(C.c_float)),
                                         r.ctypes.data_as(C.POINTER(C.c_float)) if len(r) else None, len(r), d_dst, dst_stride,
                                         C.c_void_p(stream)))

    def info(self) -> dict:
        s, n = C.c_float(), C.c_int()
        check(lib.vs_canvas_info(self._h, C.byref(s), C.byref(n)))
        return {"scale": s.value, "regions_filled": n.value}

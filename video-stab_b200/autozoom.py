"""vs::AutoZoomCrop over the C-ABI (include/vstab_b200.h, vs_auto_zoom_*): the mirror of the reference's
`AutoZoomCrop::autoZoomCrop(corrected, marginPercent)` (include/video/AutoZoomCrop.h:8-16)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._capi import check, lib

OUT_W, OUT_H = 640, 360          # hard-coded in the reference (AutoZoomCrop.cpp:246-261)


class AutoZoomCrop:
    @staticmethod
    def autoZoomCrop(corrected: np.ndarray, marginPercent: float = 0.05, device: int = 0):
        """HxWx3 uint8 BGR in; 360x640x3 out (or the frame itself when there is no content contour)."""
        if corrected is None or corrected.size == 0:
            return corrected
        if corrected.dtype != np.uint8 or corrected.ndim != 3 or corrected.shape[2] != 3:
            raise ValueError("frame must be HxWx3 uint8 (CV_8UC3 BGR)")
        f = np.ascontiguousarray(corrected)
        h, w = f.shape[:2]
        out = np.empty(max(h * w * 3, OUT_W * OUT_H * 3), np.uint8)
        ow, oh = C.c_int(), C.c_int()
        check(lib.vs_auto_zoom_crop(f.ctypes.data, w, h, f.strides[0], marginPercent, device, out.ctypes.data, 0, out.size, C.byref(ow), C.byref(oh)))
        return out[: ow.value * oh.value * 3].reshape(oh.value, ow.value, 3).copy()

    @staticmethod
    def crop_device(d_src: int, w: int, h: int, stride: int, d_out: int, out_stride: int, out_capacity: int, stream: int = 0):
        ow, oh = C.c_int(), C.c_int()
        check(lib.vs_auto_zoom_crop_device(d_src, w, h, stride, 0.05, d_out, out_stride, out_capacity, C.byref(ow), C.byref(oh), C.c_void_p(stream)))
        return ow.value, oh.value

    @staticmethod
    def rect_from_mask(mask: np.ndarray):
        """The host half on its own: (x, y, w, h) of the crop for a closed content mask, or None when there is no contour."""
        m = np.ascontiguousarray(mask, np.uint8)
        h, w = m.shape
        x, y, rw, rh, found = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(lib.vs_auto_zoom_rect_from_mask(m.ctypes.data, w, h, m.strides[0], C.byref(x), C.byref(y), C.byref(rw), C.byref(rh), C.byref(found)))
        return (x.value, y.value, rw.value, rh.value) if found.value else None

    @staticmethod
    def find_external_contours(mask: np.ndarray):
        """cv::findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) as the library follows borders (list of (n,2) int32 arrays)."""
        m = np.ascontiguousarray(mask, np.uint8)
        h, w = m.shape
        cap = 2 * h * w + 16
        pts = np.zeros(2 * cap, np.int32)
        lens = np.zeros(h * w + 16, np.int32)
        n = C.c_int()
        check(lib.vs_k_find_external_contours(m.ctypes.data, w, h, m.strides[0], pts.ctypes.data_as(C.POINTER(C.c_int)), cap,
                                              lens.ctypes.data_as(C.POINTER(C.c_int)), len(lens), C.byref(n)))
        out, k = [], 0
        for i in range(n.value):
            out.append(pts[2 * k: 2 * (k + lens[i])].reshape(-1, 2).copy())
            k += int(lens[i])
        return out

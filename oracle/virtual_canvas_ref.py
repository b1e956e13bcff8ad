"""CPU restatement of the reference's virtual-canvas output stage.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows `/root/reference/src/Stabilizer.cpp`:
    updateTemporalFrameBuffer            :2153-2167
    applyVirtualCanvasStabilization      :2066-2151
    createVirtualCanvas                  :2169-2212
    blendTemporalRegions                 :2214-2278
    calculateOptimalCanvasSize           :2280-2314
    extractTemporalRegion                :2316-2350
    seamlessBlend                        :2352-2399
    isRegionAvailable                    :2401-2421
    applyMotionCompensation              :2423-2443
Image operations are the real OpenCV of the cv2 wheel (cvtColor, threshold, findContours, boundingRect, warpAffine with
BORDER_REFLECT, resize); the float arithmetic is float32 step by step, as the C++ evaluates it.

Pinned: `tests/test_ref_pin.py::test_virtual_canvas_restatement_equals_compiled_reference` runs this file and the
compiled reference (`oracle/_ref`, the unmodified Stabilizer.cpp) on the same frames and corrections: frames bit-equal.

What the code does, in short (the names promise more): the ORIGINAL frame is pasted in the middle of a black canvas and a
frame-sized window is cut out at centre - (int)(dx, dy), i.e. the frame moves by whole pixels.  "Empty" regions are the
bounding rectangles of the EXTERNAL contours of gray <= 1; whenever the canvas is larger than the frame on all four sides the
black surround is the only external contour and the one region is the whole canvas, which an older frame covers by more than
half only for canvas scales below sqrt(2); then the most recent such frame, motion compensated and stretched over the canvas, is
alpha-blended over everything (current frame included) with weight (i + 1) / n * canvasBlendWeight.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def _rect_and(a, b):
    x1, y1 = max(a[0], b[0]), max(a[1], b[1])
    x2, y2 = min(a[0] + a[2], b[0] + b[2]), min(a[1] + a[3], b[1] + b[3])
    return (x1, y1, x2 - x1, y2 - y1) if x2 > x1 and y2 > y1 else (0, 0, 0, 0)


class VirtualCanvasRef:
    def __init__(self, canvasScaleFactor=1.5, temporalBufferSize=30, canvasBlendWeight=0.7, adaptiveCanvasSize=True,
                 maxCanvasScale=2.0, minCanvasScale=1.2, edgeBlendRadius=20, **_):
        import cv2
        self.cv2 = cv2
        self.scale0 = f32(canvasScaleFactor)
        self.tbuf = int(temporalBufferSize)
        self.blend_w = f32(canvasBlendWeight)
        self.adaptive = bool(adaptiveCanvasSize)
        self.max_s, self.min_s = f32(maxCanvasScale), f32(minCanvasScale)
        self.edge = int(edgeBlendRadius)
        self.scale = f32(canvasScaleFactor)
        self.canvas = None
        self.frames: list[np.ndarray] = []
        self.corrections: list[np.ndarray] = []
        self.filled: list[tuple] = []

    def _optimal_scale(self, transforms):                                       # :2280-2314
        mx = f32(0)
        for m in transforms[-30:]:
            mx = max(mx, np.sqrt(f32(f32(m[0] * m[0]) + f32(m[1] * m[1]))))
        factor = max(f32(1.0), f32(mx / f32(50.0)))
        opt = f32(self.scale0 + f32(f32(factor - f32(1.0)) * f32(0.5)))
        return max(self.min_s, min(self.max_s, opt))

    def apply(self, frame: np.ndarray, correction, transforms=()) -> np.ndarray:
        """`frame` with its correction (dx, dy, da); `transforms` = transforms_ (read when the canvas is first sized)."""
        cv2 = self.cv2
        T = np.asarray(correction, f32)
        self.frames.append(frame.copy())                                        # :2153-2167
        self.corrections.append(T.copy())
        while len(self.frames) > self.tbuf:
            self.frames.pop(0)
            self.corrections.pop(0)
        h, w = frame.shape[:2]
        if self.canvas is None or self.canvas[0] != int(f32(w) * self.scale) or self.canvas[1] != int(f32(h) * self.scale):
            self.scale = self._optimal_scale(transforms) if (self.adaptive and len(transforms)) else self.scale0
            self.canvas = (int(f32(w) * self.scale), int(f32(h) * self.scale))
            self.center = (f32(self.canvas[0]) / f32(2.0), f32(self.canvas[1]) / f32(2.0))
        cw, ch = self.canvas
        ox, oy = f32(self.center[0] - f32(w) / f32(2.0)), f32(self.center[1] - f32(h) / f32(2.0))
        res = np.zeros((ch, cw, 3), np.uint8)                                   # :2169-2212
        fr = (int(ox), int(oy), w, h)
        v = _rect_and(fr, (0, 0, cw, ch))
        if v[2] > 0 and v[3] > 0:
            sx, sy = v[0] - fr[0], v[1] - fr[1]
            res[v[1]:v[1] + v[3], v[0]:v[0] + v[2]] = frame[sy:sy + v[3], sx:sx + v[2]]
        self.filled = []
        n = len(self.frames)
        if n >= 2:                                                              # :2214-2278
            gray = cv2.cvtColor(res, cv2.COLOR_BGR2GRAY)
            _, binary = cv2.threshold(gray, 1, 255, cv2.THRESH_BINARY_INV)
            contours, _ = cv2.findContours(binary, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
            for r in (cv2.boundingRect(c) for c in contours):
                if r[2] * r[3] <= 100:
                    continue
                best, best_w = None, f32(0)
                for i in range(n - 1):
                    rel = f32(T - self.corrections[i])
                    fh, fw = self.frames[i].shape[:2]
                    inter = _rect_and((r[0] + int(rel[0]), r[1] + int(rel[1]), r[2], r[3]), (0, 0, fw, fh))
                    if not f32(f32(inter[2] * inter[3]) / f32(r[2] * r[3])) > f32(0.5):       # :2401-2421
                        continue
                    tw = f32(f32(i + 1) / f32(n)) * self.blend_w
                    if tw > best_w:
                        best, best_w = (i, rel, inter), tw
                if best is None:
                    continue
                i, rel, inter = best
                da = -rel[2]                                                    # :2423-2443
                M = np.array([[np.cos(da, dtype=f32), -np.sin(da, dtype=f32), -rel[0]],
                              [np.sin(da, dtype=f32), np.cos(da, dtype=f32), -rel[1]]], f32)
                src = self.frames[i]
                comp = cv2.warpAffine(src, M, (src.shape[1], src.shape[0]), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
                fill = comp[inter[1]:inter[1] + inter[3], inter[0]:inter[0] + inter[2]]
                if (fill.shape[1], fill.shape[0]) != (r[2], r[3]):              # :2341-2344
                    fill = cv2.resize(fill, (r[2], r[3]), interpolation=cv2.INTER_LINEAR)
                self.filled.append((r, i, float(best_w)))
                edge = min(self.edge, min(r[2], r[3]) // 4)                     # :2352-2399
                yy, xx = np.mgrid[0:r[3], 0:r[2]]
                d = np.minimum(np.minimum(xx, yy), np.minimum(r[2] - xx - 1, r[3] - yy - 1)).astype(f32)
                alpha = np.full((r[3], r[2]), best_w, f32)
                if edge > 0:
                    alpha = np.where(d < edge, (alpha * (d / f32(edge)).astype(f32)).astype(f32), alpha)
                a = alpha[..., None]
                tgt = res[r[1]:r[1] + r[3], r[0]:r[0] + r[2]]
                val = ((f32(1.0) - a).astype(f32) * tgt.astype(f32)).astype(f32) + (a * fill.astype(f32)).astype(f32)
                tgt[...] = val.astype(f32).astype(np.uint8)
        fox, foy = f32(ox - T[0]), f32(oy - T[1])                               # :2116-2147
        ex, ey, ew, eh = max(0, int(fox)), max(0, int(foy)), w, h
        ex, ey = min(ex, cw - ew), min(ey, ch - eh)
        ew, eh = min(ew, cw - ex), min(eh, ch - ey)
        if ew > 0 and eh > 0 and ex >= 0 and ey >= 0 and ex + ew <= cw and ey + eh <= ch:
            out = res[ey:ey + eh, ex:ex + ew].copy()
            if (out.shape[1], out.shape[0]) != (w, h):
                out = cv2.resize(out, (w, h), interpolation=cv2.INTER_LANCZOS4)
            return out
        return frame

"""Python restatement of the reference's CPU stabilizer (`useCuda=false` branch).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows, statement by statement,
`/root/reference/src/Stabilizer.cpp`:

    ctor                         :50-164
    stabilize / flush / clean    :258-400, :221-256
    generateTransform (CPU)      :402-411, :446-456, :594-678, :680-693, :695-697, :737-746, :756-760
    applyNextSmoothTransform     :763-908, :979-991, :1047-1067, :1107-1137
    boxFilterConvolve            :1139-1172
    gaussianFilterConvolve       :1364-1413       kalmanFilterSmooth   :1416-1458
    adaptSmoothingRadius         :1461-1492       updateAdaptiveParameters :1562-1574
    calculateAdaptiveRadius      :1637-1673       analyzeMotionIntent  :1676-1719
    calculateAdaptiveStabilizationStrength :1722-1747   variance/consistency :1750-1780

calling the real OpenCV (Python `cv2` 4.13.0) for the library operations.  All host
arithmetic the reference does in `float` is done in numpy float32 scalars in the same
order; `std::cos/sin/atan2/sqrt/exp(float)` go to glibc's `cosf/sinf/atan2f/sqrtf/expf`
through ctypes, exactly what the C++ would call on this box.

`border_type: fade` (:914-978, :1070-1106) is restated too, including the reference's mask quirk (the second
cv::rectangle paints the WHOLE mask 255, so the history is blended into and updated from the whole frame, not just
the border band).  cv::addWeighted is the real library call with optimisations ON (its SIMD path fuses
`src1*alpha + (src2*beta)`; the plain path differs by <= 1 LSB); the history update `(1-0.1f)*h + 0.1f*s` is
evaluated in float32 without contraction.

`drone_high_freq_mode` (:2447-2686) is restated for the case its analysis size comes out as 960x540 (the default
`hfAnalysisMaxWidth` = 960 on any 16:9 frame): dead-zone freeze, micro-shake suppression, rotation low-pass (only with
`horizonLock`), the 10-sample translation history, and the [10,50] box-radius clamp.  Conditional CLAHE never fires in
the reference (both call sites pass -1 to shouldApplyConditionalCLAHE, which then always returns false).  The HF state is
set in the constructor only; clean() does not reset it.

`enable_virtual_canvas` is restated separately, in oracle/virtual_canvas_ref.py.

The two process-global `static` counters of the reference (`frameTicker` :260,
`featureDetectionCounter` :696) are per-instance here: parity is defined per stream
against a fresh single-instance reference run.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import dataclasses
from collections import deque

import numpy as np

f32 = np.float32

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("cosf", "sinf", "sqrtf", "expf"):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float]
_libm.atan2f.restype = ctypes.c_float
_libm.atan2f.argtypes = [ctypes.c_float, ctypes.c_float]


def cosf(x): return f32(_libm.cosf(float(x)))
def sinf(x): return f32(_libm.sinf(float(x)))
def sqrtf(x): return f32(_libm.sqrtf(float(x)))
def expf(x): return f32(_libm.expf(float(x)))
def atan2f(y, x): return f32(_libm.atan2f(float(y), float(x)))


@dataclasses.dataclass
class Parameters:
    """vs::Stabilizer::Parameters, Stabilizer.h:76-175 (live fields only; the inert ones are
    accepted by the product's `vs_params` and ignored)."""
    useCuda: bool = False
    logging: bool = False
    smoothingRadius: int = 30
    maxCorners: int = 200
    qualityLevel: float = 0.01
    minDistance: float = 30.0
    blockSize: int = 3
    borderType: str = "black"
    borderSize: int = 0
    cropNZoom: bool = False
    smoothingMethod: str = "box"
    gaussianSigma: float = 2.0
    horizonLock: bool = False
    adaptiveSmoothing: bool = False
    minSmoothingRadius: int = 5
    maxSmoothingRadius: int = 50
    fadeAlpha: float = 0.1
    fadeDuration: int = 30
    droneHighFreqMode: bool = False
    hfShakePx: float = 1.5
    hfAnalysisMaxWidth: int = 960
    hfRotLPAlpha: float = 0.2
    enableConditionalCLAHE: bool = True
    hfDeadZoneThreshold: float = 2.0
    hfFreezeDuration: int = 10
    hfMotionAccumulatorDecay: float = 0.9


_BORDER = {"reflect": 2, "reflect_101": 4, "replicate": 1, "wrap": 3}     # mapBorderMode :31-38

INTENT_NORMAL, INTENT_PAN, INTENT_SHAKE, INTENT_FOLLOW = 0, 1, 2, 3


@dataclasses.dataclass
class FrameRecord:
    """Everything the reference computes for one generateTransform() call."""
    frame_index: int
    prev_pts: np.ndarray            # keypoints LK started from (N,2) f32
    next_pts: np.ndarray            # LK output (N,2) f32
    status: np.ndarray              # (N,) u8
    inlier_mask: np.ndarray | None  # over the status-filtered pairs
    affine: np.ndarray | None       # 2x3 f64 from estimateAffinePartial2D
    transform: np.ndarray           # (dx,dy,da) f32
    path: np.ndarray                # cumulative f32
    detected: np.ndarray | None     # corners re-detected on this frame (or None)


@dataclasses.dataclass
class OutputRecord:
    index: int
    path_len: int
    radius: int
    intent: int
    smoothed: np.ndarray            # smoothedPath_[index]
    T: np.ndarray | None            # 2x3 f32 (None => passthrough)


class StabilizerRef:
    def __init__(self, params: Parameters, use_optimized: bool = False):
        import cv2
        self.cv2 = cv2
        cv2.setUseOptimized(use_optimized)
        self.p = dataclasses.replace(params)
        self.border_mode = _BORDER.get(self.p.borderType, 0)
        if self.p.cropNZoom and self.p.borderType != "black":
            self.border_mode = 0
        self.border_history = None                               # borderHistory_ / fadeFrameCount_ (:73-77): NOT reset by clean()
        self.fade_count = 0
        # drone high-frequency state (:143-153), constructor only
        self.hf_hist: list = []
        self.hf_median = np.zeros(2, f32)
        self.hf_rot_lp = f32(0)
        self.hf_in_dead_zone = False
        self.hf_freeze_counter = 0
        self.hf_accumulator = f32(0)
        self.frame_records: list[FrameRecord] = []
        self.output_records: list[OutputRecord] = []
        self.clean()

    # ---------------------------------------------------------------- :221-256
    def clean(self):
        self.queue: deque = deque()
        self.index_queue: deque = deque()
        self.transforms: list[np.ndarray] = []
        self.path: list[np.ndarray] = []
        self.prev_gray = None
        self.prev_kp = np.zeros((0, 2), f32)
        self.first = True
        self.next_index = 0
        self.orig_size = None
        self.detect_counter = 0

    # ---------------------------------------------------------------- :258-392
    def stabilize(self, frame: np.ndarray):
        cv2 = self.cv2
        if frame is None or frame.size == 0:
            return None
        if self.p.cropNZoom and self.orig_size is None:
            self.orig_size = (frame.shape[1], frame.shape[0])
        if self.first:
            small = cv2.resize(frame, (480, 270), interpolation=cv2.INTER_LINEAR)
            self.prev_gray = cv2.cvtColor(small, cv2.COLOR_BGR2GRAY)
            c = cv2.goodFeaturesToTrack(self.prev_gray, self.p.maxCorners, self.p.qualityLevel,
                                        self.p.minDistance, None, blockSize=self.p.blockSize)
            self.prev_kp = np.zeros((0, 2), f32) if c is None else c.reshape(-1, 2).copy()
            self.first_corners = self.prev_kp.copy()
            self.queue.append(frame.copy())
            self.index_queue.append(0)
            self.first = False
            self.next_index = 1
            return None
        self.queue.append(frame)
        self.index_queue.append(self.next_index)
        self._generate_transform(frame)
        gate = max(5, min(self.p.smoothingRadius, 35))
        if len(self.index_queue) < gate:
            self.next_index += 1
            return None
        out = self._apply_next()
        self.next_index += 1
        return out

    def flush(self):                                              # :394-400
        if not self.queue:
            return None
        return self._apply_next()

    # ---------------------------------------------------------------- :402-761
    def _generate_transform(self, frame):
        cv2 = self.cv2
        small = cv2.resize(frame, (960, 540), interpolation=cv2.INTER_LINEAR)
        gray = cv2.cvtColor(small, cv2.COLOR_BGR2GRAY)
        rec_prev = self.prev_kp
        nxt = np.zeros((0, 2), f32)
        status = np.zeros((0,), np.uint8)
        mask = None
        affine = None
        if len(self.prev_kp) and self.prev_gray is not None:
            if self.prev_gray.shape != gray.shape:                 # :598-603 (fires on frame 1)
                self.prev_gray = cv2.resize(self.prev_gray, (960, 540), interpolation=cv2.INTER_LINEAR)
            nxt, st, _ = cv2.calcOpticalFlowPyrLK(
                self.prev_gray, gray, self.prev_kp, None, winSize=(15, 15), maxLevel=2,
                criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 20, 0.03))
            nxt = nxt.reshape(-1, 2)
            status = st.ravel()
            ok = status != 0
            vp, vc = self.prev_kp[ok], nxt[ok]
            t = np.array([[1, 0, 0], [0, 1, 0]], f32)
            if len(vp) >= 4:
                a, m = cv2.estimateAffinePartial2D(vp, vc, None, cv2.RANSAC, 5.0, 500)
                if a is not None and a.shape == (2, 3):
                    affine = a
                    mask = m.ravel()
                    t = a.astype(f32)
            dx, dy = t[0, 2], t[1, 2]
            da = atan2f(t[1, 0], t[0, 0])
            tr = np.array([dx, dy, da], f32)
            if self.p.droneHighFreqMode:                           # :666-671
                tr = self._hf_filters(tr)
        else:
            tr = np.zeros(3, f32)                                  # :675-677
        self.transforms.append(tr)
        self.path.append(tr.copy() if not self.path else (self.path[-1] + tr).astype(f32))
        if self.p.adaptiveSmoothing:                               # :691-693
            self._update_adaptive()
        detected = None
        self.detect_counter += 1
        if self.detect_counter % 2 == 0:                           # :696-746
            c = cv2.goodFeaturesToTrack(gray, min(self.p.maxCorners, 200), 0.02, 15.0, None, blockSize=3)
            self.prev_kp = np.zeros((0, 2), f32) if c is None else c.reshape(-1, 2).copy()
            detected = self.prev_kp.copy()
        self.prev_gray = gray
        self.frame_records.append(FrameRecord(len(self.transforms), rec_prev, nxt, status, mask, affine,
                                              tr, self.path[-1].copy(), detected))

    # --------------------------------------------------------- :2468-2686 (drone high-frequency chain)
    def _hf_filters(self, raw):
        p = self.p
        dx, dy, da = f32(raw[0]), f32(raw[1]), f32(raw[2])
        thr = f32(p.hfDeadZoneThreshold)
        # applyDeadZoneFreeze :2605-2655 (updateMotionAccumulator :2668-2681 first)
        mag = sqrtf(f32(f32(dx * dx) + f32(dy * dy)) + f32(f32(da * da) * f32(100.0)))
        acc = max(f32(self.hf_accumulator * f32(p.hfMotionAccumulatorDecay)), mag)
        acc = min(acc, f32(thr * f32(5.0)))
        acc = max(f32(0), min(acc, f32(100.0)))
        self.hf_accumulator = f32(acc)
        out = np.array([dx, dy, da], f32)
        if not self.hf_in_dead_zone and mag < thr:
            self.hf_in_dead_zone = True
            self.hf_freeze_counter = p.hfFreezeDuration
        if self.hf_in_dead_zone:
            self.hf_freeze_counter -= 1
            if (self.hf_freeze_counter <= 0 or mag > f32(thr * f32(1.5)) or self.hf_accumulator > f32(thr * f32(1.2))):
                self.hf_in_dead_zone = False
                self.hf_freeze_counter = 0
                self.hf_accumulator = f32(0)
            else:
                out = np.zeros(3, f32)
        # applyMicroShakeSuppression :2468-2503
        if len(self.hf_hist) >= 5:
            xs = sorted(f32(t[0]) for t in self.hf_hist)
            ys = sorted(f32(t[1]) for t in self.hf_hist)
            mid = len(xs) // 2
            if len(xs) % 2 == 0:
                self.hf_median = np.array([f32(f32(xs[mid - 1] + xs[mid]) / f32(2)), f32(f32(ys[mid - 1] + ys[mid]) / f32(2))], f32)
            else:
                self.hf_median = np.array([xs[mid], ys[mid]], f32)
        dev = (out[:2] - self.hf_median).astype(f32)
        m2 = sqrtf(f32(f32(dev[0] * dev[0]) + f32(dev[1] * dev[1])))
        shake = f32(p.hfShakePx)
        if m2 < shake:
            out[:2] = (self.hf_median + (dev * f32(0.01)).astype(f32)).astype(f32)
        elif m2 < f32(shake * f32(2.0)):
            out[:2] = (self.hf_median + (dev * f32(0.05)).astype(f32)).astype(f32)
        # applyRotationLowPass :2505-2520
        if p.horizonLock:
            a = f32(p.hfRotLPAlpha)
            self.hf_rot_lp = f32(f32(f32(f32(1.0) - a) * self.hf_rot_lp) + f32(a * out[2]))
            out[2] = self.hf_rot_lp
        # updateTranslationHistory :2522-2529
        self.hf_hist.append(out[:2].copy())
        if len(self.hf_hist) > 10:
            self.hf_hist.pop(0)
        return out

    # --------------------------------------------------------- :1461-1492,1562-1574
    def _update_adaptive(self):
        if len(self.transforms) < 3:
            return
        m = self.transforms[-1]
        mag = sqrtf(m[0] * m[0] + m[1] * m[1])
        scale = max(f32(0), min(f32(1), mag / f32(50)))
        scale = f32(1) - scale
        new_r = self.p.minSmoothingRadius + int(scale * f32(self.p.maxSmoothingRadius - self.p.minSmoothingRadius))
        if new_r != self.p.smoothingRadius:
            self.p.smoothingRadius = new_r

    # --------------------------------------------------------------- :1139-1172
    @staticmethod
    def _box(path: np.ndarray, radius: int, drone: bool = False) -> np.ndarray:
        r = max(10, min(radius, 50)) if drone else max(2, min(radius, 8))
        n = len(path)
        if n <= r:
            return path.copy()
        out = np.empty(n, f32)
        for i in range(n):
            s = f32(0)
            lo, hi = max(0, i - r), min(n - 1, i + r)
            for j in range(lo, hi + 1):
                s = s + path[j]
            out[i] = s / f32(hi - lo + 1)
        return out

    @staticmethod
    def box_at(path: np.ndarray, radius: int, i: int, drone: bool = False) -> np.float32:
        """smoothed[i] only (what the reference actually consumes)."""
        r = max(10, min(radius, 50)) if drone else max(2, min(radius, 8))
        n = len(path)
        if n <= r:
            return path[i]
        s = f32(0)
        lo, hi = max(0, i - r), min(n - 1, i + r)
        for j in range(lo, hi + 1):
            s = s + path[j]
        return s / f32(hi - lo + 1)

    # --------------------------------------------------------------- :1364-1413
    @staticmethod
    def _gaussian(path: np.ndarray, sigma_d: float) -> np.ndarray:
        sigma = f32(sigma_d)                                        # float parameter
        ksz = max(3, int(np.ceil(f32(6) * sigma)))
        if ksz % 2 == 0:
            ksz += 1
        c = ksz // 2
        kern = np.empty(ksz, f32)
        tot = f32(0)
        for i in range(ksz):
            x = f32(i - c)
            kern[i] = expf(-(x * x) / (f32(2) * sigma * sigma))
            tot = tot + kern[i]
        kern = (kern / tot).astype(f32)
        n = len(path)
        if n <= c:
            # B-Q8: the reference reads out of bounds here (UB).  Defined behaviour for the
            # drop-in: fall back to the box filter until the path is longer than `center`.
            return None
        pad = np.empty(n + 2 * c, f32)
        for i in range(c):
            pad[i] = path[c - i]
        pad[c:c + n] = path
        for i in range(c):
            pad[c + n + i] = path[n - 1 - i]
        out = np.empty(n, f32)
        for i in range(n):
            s = f32(0)
            for j in range(ksz):
                s = s + pad[i + j] * kern[j]
            out[i] = s
        return out

    # --------------------------------------------------------------- :1416-1458
    def _kalman(self, path: np.ndarray) -> np.ndarray:
        cv2 = self.cv2
        kf = cv2.KalmanFilter(2, 1, 0)
        kf.transitionMatrix = np.array([[1, 1], [0, 1]], f32)
        kf.measurementMatrix = np.array([[1, 0]], f32)
        kf.processNoiseCov = np.array([[0.01, 0], [0, 0.01]], f32)
        kf.measurementNoiseCov = np.array([[0.1]], f32)
        kf.statePost = np.array([[path[0]], [0]], f32)
        out = np.empty(len(path), f32)
        out[0] = path[0]
        for i in range(1, len(path)):
            kf.predict()
            out[i] = kf.correct(np.array([[path[i]]], f32))[0, 0]
        return out

    # --------------------------------------------------------------- :1637-1673
    def _adaptive_radius(self, px, py, pa) -> int:
        n = len(px)
        if n < 10:
            return self.p.smoothingRadius
        start = max(0, n - 20)
        cnt = f32(n - start)
        mx = my = ma = f32(0)
        for i in range(start, n):
            mx = mx + px[i]
            my = my + py[i]
            ma = ma + pa[i]
        mx, my, ma = mx / cnt, my / cnt, ma / cnt
        vx = vy = va = f32(0)
        for i in range(start, n):
            vx = vx + (px[i] - mx) * (px[i] - mx)
            vy = vy + (py[i] - my) * (py[i] - my)
            va = va + (pa[i] - ma) * (pa[i] - ma)
        vx, vy, va = vx / cnt, vy / cnt, va / cnt
        total = sqrtf(vx + vy + va * f32(1000))
        return int(max(f32(5), min(f32(25), total * f32(2))))

    # --------------------------------------------------------------- :1750-1780
    @staticmethod
    def _variance(v) -> np.float32:
        if not len(v):
            return f32(0)
        m = f32(0)
        for x in v:
            m = m + x
        m = m / f32(len(v))
        var = f32(0)
        for x in v:
            d = x - m
            var = var + d * d
        return var / f32(len(v))

    @classmethod
    def _consistency(cls, v) -> np.float32:
        if len(v) < 2:
            return f32(0)
        var = cls._variance(v)
        m = f32(0)
        for x in v:
            m = m + x
        m = m / f32(len(v))
        if m == 0:
            return f32(0)
        c = f32(1) / (f32(1) + (var / (m * m)))
        return max(f32(0), min(f32(1), c))

    # --------------------------------------------------------------- :1676-1719
    def _intent(self, motion, idx: int) -> int:
        mag = sqrtf(motion[0] * motion[0] + motion[1] * motion[1])
        ang = f32(float(abs(motion[2]) * f32(180)) / np.pi * float(f32(30)))
        if len(self.transforms) >= 15:
            mags, dirs = [], []
            for i in range(max(0, idx - 15), idx):
                if i < len(self.transforms):
                    t = self.transforms[i]
                    mags.append(sqrtf(t[0] * t[0] + t[1] * t[1]))
                    dirs.append(atan2f(t[1], t[0]))
            if mags:
                dv = self._variance(dirs)
                mc = self._consistency(mags)
                if dv < f32(0.5) and mc > f32(0.7) and mag > f32(5):
                    return INTENT_PAN
                if mag < f32(3) and mc < f32(0.3) and ang > f32(10):
                    return INTENT_SHAKE
                if mag > f32(3) and mag < f32(15) and dv > f32(0.5):
                    return INTENT_FOLLOW
        return INTENT_NORMAL

    # ---------------------------------------------------------------- :763-1137
    def _apply_next(self):
        cv2 = self.cv2
        if not self.queue:
            return None
        frame = self.queue.popleft()
        idx = self.index_queue.popleft()
        if idx >= len(self.transforms):                            # :774-780  passthrough
            self.output_records.append(OutputRecord(idx, len(self.path), 0, 0, np.zeros(3, f32), None))
            return frame
        arr = np.array(self.path, f32)
        px, py, pa = arr[:, 0].copy(), arr[:, 1].copy(), arr[:, 2].copy()
        radius = 0
        sm = None
        if self.p.smoothingMethod == "gaussian":
            g = [self._gaussian(c, self.p.gaussianSigma) for c in (px, py, pa)]
            if g[0] is not None:
                sm = np.array([g[0][idx], g[1][idx], g[2][idx]], f32)
        elif self.p.smoothingMethod == "kalman":
            sm = np.array([self._kalman(c)[idx] for c in (px, py, pa)], f32)
        if sm is None:
            radius = self._adaptive_radius(px, py, pa)
            sm = np.array([self.box_at(c, radius, idx, self.p.droneHighFreqMode) for c in (px, py, pa)], f32)
        raw = self.transforms[idx]
        diff = (sm - arr[idx]).astype(f32)
        intent = INTENT_NORMAL
        if idx > 0:                                                # :854-888
            intent = self._intent(raw, idx)
            if intent == INTENT_PAN:
                diff = diff * f32(0.5)
            elif intent == INTENT_SHAKE:
                diff = diff * f32(1.0)
            elif intent == INTENT_FOLLOW:
                diff = diff * f32(0.8)
            else:
                diff = diff * f32(0.7)                             # calculateAdaptive...Strength default
        ts = (raw + diff).astype(f32)
        dx, dy, da = ts
        if self.p.horizonLock:
            da = f32(0)
        T = np.array([[cosf(da), -sinf(da), dx], [sinf(da), cosf(da), dy]], f32)
        self.output_records.append(OutputRecord(idx, len(self.path), radius, intent, sm, T))
        src = frame
        fade = self.p.borderType == "fade"
        if fade:                                                   # :914-978
            if self.p.borderSize > 0 and not self.p.cropNZoom:
                b = self.p.borderSize
                if self.border_history is None:
                    self.border_history = cv2.copyMakeBorder(frame, b, b, b, b, cv2.BORDER_CONSTANT, value=(0, 0, 0))
                    self.fade_count = 0
                src = cv2.copyMakeBorder(frame, b, b, b, b, cv2.BORDER_CONSTANT, value=(0, 0, 0))
                alpha = f32(self.p.fadeAlpha)
                if self.fade_count < self.p.fadeDuration:
                    alpha = f32(alpha * f32(f32(self.fade_count) / f32(self.p.fadeDuration)))
                    self.fade_count += 1
                beta = f32(f32(1.0) - alpha)
                if self.border_history.shape == src.shape:
                    opt = cv2.useOptimized()
                    cv2.setUseOptimized(True)
                    src = cv2.addWeighted(self.border_history, float(alpha), src, float(beta), 0.0)   # mask is all 255
                    cv2.setUseOptimized(opt)
        elif self.p.borderSize > 0 and not self.p.cropNZoom:       # :981-990
            b = self.p.borderSize
            src = cv2.copyMakeBorder(frame, b, b, b, b, self.border_mode, value=(0, 0, 0))
        out = cv2.warpAffine(src, T, (src.shape[1], src.shape[0]), flags=cv2.INTER_LINEAR,
                             borderMode=cv2.BORDER_CONSTANT)
        if fade and self.p.borderSize > 0:                         # :1070-1106
            if self.border_history is not None and self.border_history.shape == out.shape:
                h = self.border_history.astype(f32)
                upd = (f32(f32(1.0) - f32(0.1)) * h).astype(f32) + (f32(0.1) * out.astype(f32)).astype(f32)
                self.border_history = upd.astype(f32).astype(np.uint8)      # static_cast<uchar>: truncation
            else:
                self.border_history = out.copy()
        if self.p.cropNZoom and self.p.borderSize > 0:             # :1108-1124
            b = self.p.borderSize
            w, h = out.shape[1] - 2 * b, out.shape[0] - 2 * b
            if w <= 0 or h <= 0:
                return out
            crop = out[b:b + h, b:b + w].copy()
            if self.orig_size is not None:
                crop = cv2.resize(crop, self.orig_size)
            return crop
        return out


def run_clip(frames, params: Parameters, flush: bool = True, use_optimized: bool = False):
    """Push every frame, then flush.  Returns (outputs list aligned with pops, StabilizerRef)."""
    st = StabilizerRef(params, use_optimized=use_optimized)
    outs = []
    for f in frames:
        o = st.stabilize(f)
        if o is not None:
            outs.append(o)
    if flush:
        while True:
            o = st.flush()
            if o is None:
                break
            outs.append(o)
    return outs, st

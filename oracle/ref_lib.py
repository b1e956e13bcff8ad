"""Driver for oracle/_ref/libvideostab_ref.so — the reference's own `Stabilizer.cpp`, compiled unmodified against the
mini_cv header stand-in (see oracle/build_ref.py).  TEST INFRASTRUCTURE ONLY.

This module supplies the other half: the `mini_cv_ops` callback table.  Every callback is the REAL OpenCV
function of the cv2 4.13 wheel — `cv::resize` -> `cv2.resize`, `cv::calcOpticalFlowPyrLK` -> `cv2.calcOpticalFlowPyrLK`,
`cv::KalmanFilter::predict/correct` -> a `cv2.KalmanFilter` object, and so on.  Net effect: the reference's host
logic is the reference's C++ and the library arithmetic is the library.  `RefStabilizer` mirrors
`vs::Stabilizer` (`include/video/Stabilizer.h:177-198`: ctor(params), stabilize, flush, clean) and records, per
`generateTransform()` call, what went through the callbacks (LK points + status, RANSAC mask, detected corners),
so tests can compare the same quantities they compare for the Python restatement and for the CUDA path.

The reference keeps two process-global function-local statics (`frameTicker` Stabilizer.cpp:260,
`featureDetectionCounter` :696).  Parity is defined against a FRESH single-instance run, so every
`RefStabilizer` loads its own private copy of the shared library (fresh statics).
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import tempfile

import numpy as np

from .stabilizer_ref import FrameRecord, OutputRecord

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libvideostab_ref.so")
f32 = np.float32


def available() -> bool:
    """True when the compiled reference is on disk (building it first when /root/reference is here)."""
    try:
        from . import build_ref
        build_ref.build()
    except Exception:
        pass
    return os.path.exists(LIB)


# ---------------------------------------------------------------------------------------------- callback table
class MiniMat(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_ubyte)), ("rows", C.c_int), ("cols", C.c_int), ("type", C.c_int), ("step", C.c_size_t)]


_DEPTH = {0: np.uint8, 1: np.int8, 2: np.uint16, 3: np.int16, 4: np.int32, 5: np.float32, 6: np.float64}
PM = C.POINTER(MiniMat)
PF = C.POINTER(C.c_float)
PD = C.POINTER(C.c_double)
PU = C.POINTER(C.c_ubyte)
PI = C.POINTER(C.c_int)


def _np(m) -> np.ndarray:
    """numpy view (no copy) of a mini_cv_mat."""
    m = m.contents if hasattr(m, "contents") else m
    dt = np.dtype(_DEPTH[m.type & 7])
    cn = (m.type >> 3) + 1
    if m.rows == 0 or m.cols == 0 or not m.data:
        return np.zeros((0, 0, cn) if cn > 1 else (0, 0), dt)
    nbytes = (m.rows - 1) * m.step + m.cols * cn * dt.itemsize
    buf = (C.c_ubyte * nbytes).from_address(C.addressof(m.data.contents))
    if cn > 1:
        return np.ndarray((m.rows, m.cols, cn), dt, buffer=buf, strides=(m.step, cn * dt.itemsize, dt.itemsize))
    return np.ndarray((m.rows, m.cols), dt, buffer=buf, strides=(m.step, dt.itemsize))


def _put(dst, arr):
    out = _np(dst)
    if arr.ndim == 2 and out.ndim == 3:
        arr = arr[:, :, None]
    if out.shape != arr.shape:
        raise ValueError(f"mini_cv: destination {out.shape} != result {arr.shape}")
    if arr.ctypes.data != out.ctypes.data:        # cv2 was handed `out` as dst= and wrote in place: nothing to copy
        out[...] = arr


def _dst(m):
    """the destination Mat as a cv2 `dst=` argument (cv2 writes in place when shape, type and contiguity fit)"""
    out = _np(m)
    return out if out.flags.c_contiguous else None


_SIG = [
    ("resize", [PM, PM, C.c_int]),
    ("cvt_color", [PM, PM, C.c_int]),
    ("gftt", [PM, PM, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, PF, C.c_int, PI]),
    ("pyr_lk", [PM, PM, PF, C.c_int, PF, PU, PF, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double]),
    ("estimate_affine_partial", [PF, PF, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int, PD, PU, PI]),
    ("warp_affine", [PM, PM, PD, C.c_int, C.c_int, C.c_int, PD]),
    ("copy_make_border", [PM, PM, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, PD]),
    ("add_weighted", [PM, C.c_double, PM, C.c_double, C.c_double, PM]),
    ("threshold", [PM, PM, C.c_double, C.c_double, C.c_int]),
    ("find_contours", [PM, C.c_int, C.c_int, PI, C.c_int, PI, C.c_int, PI]),
    ("kalman_create", [C.c_int, C.c_int, C.c_int, PI]),
    ("kalman_predict", [C.c_int, PM]),
    ("kalman_correct", [C.c_int, PM, PM]),
    ("kalman_release", [C.c_int]),
    ("canny", [PM, PM, C.c_double, C.c_double, C.c_int, C.c_int]),
    ("hough_lines", [PM, C.c_double, C.c_double, C.c_int, C.c_int, PF, PI, C.c_int, PI]),
    ("gaussian_blur", [PM, PM, C.c_int, C.c_int, C.c_double, C.c_double]),
    ("remap", [PM, PM, PM, PM, C.c_int, C.c_int]),
    ("morphology", [PM, PM, C.c_int, PM]),
    ("structuring_element", [C.c_int, PM]),
    ("draw_contours", [PM, PI, PI, C.c_int, C.c_int, PD, C.c_int]),
    ("rotation_matrix", [C.c_double, C.c_double, C.c_double, C.c_double, PD]),
    ("sobel", [PM, PM, C.c_int, C.c_int, C.c_int]),
]
_FN = {name: C.CFUNCTYPE(C.c_int, *args) for name, args in _SIG}


class OpsTable(C.Structure):
    _fields_ = [(name, _FN[name]) for name, _ in _SIG]


class CvOps:
    """The callbacks.  `log` (when set) receives (op name, dict) for the calls tests want to see."""

    def __init__(self, use_optimized: bool = False):
        import cv2
        self.cv2 = cv2
        self.use_optimized = use_optimized
        self.log = None
        self.errors: list[str] = []
        self._kalman: dict[int, object] = {}
        self._next_handle = 1
        self.seconds_in_cv = 0.0
        fns = {}
        for name, _ in _SIG:
            fns[name] = _FN[name](self._guard(getattr(self, "_" + name), name))
        self._keep = fns
        self.table = OpsTable(**fns)

    def _guard(self, fn, name):
        def call(*a):
            try:
                self.cv2.setUseOptimized(self.use_optimized)
                fn(*a)
                return 0
            except Exception as e:   # cv2.error -> non-zero -> mini_cv throws cv::Exception, as the real library would
                self.errors.append(f"{name}: {e}")
                return 1
        return call

    def _emit(self, op, **kw):
        if self.log is not None:
            self.log(op, kw)

    # ---- the operations on the stabilize() path
    def _resize(self, src, dst, interp):
        d = _np(dst)
        _put(dst, self.cv2.resize(_np(src), (d.shape[1], d.shape[0]), dst=_dst(dst), interpolation=interp))

    def _cvt_color(self, src, dst, code):
        _put(dst, self.cv2.cvtColor(_np(src), code, dst=_dst(dst)))

    def _gftt(self, img, mask, max_corners, quality, min_dist, block, harris, k, xy, cap, n):
        m = _np(mask) if mask else None
        c = self.cv2.goodFeaturesToTrack(_np(img), max_corners, quality, min_dist, mask=m, blockSize=block,
                                         useHarrisDetector=bool(harris), k=k)
        pts = np.zeros((0, 2), f32) if c is None else c.reshape(-1, 2).astype(f32)
        if len(pts) > cap:
            raise ValueError("gftt: too many corners for the buffer")
        n[0] = len(pts)
        if len(pts):
            np.ctypeslib.as_array(xy, shape=(len(pts) * 2,))[:] = pts.ravel()
        self._emit("gftt", corners=pts.copy(), shape=_np(img).shape, args=(max_corners, quality, min_dist, block))

    def _pyr_lk(self, prev, nxt, prev_xy, n, next_xy, status, err, ww, wh, max_level, ctype, ccount, ceps, flags, min_eig):
        p0 = np.ctypeslib.as_array(prev_xy, shape=(n * 2,)).reshape(n, 2).astype(f32)
        p1, st, er = self.cv2.calcOpticalFlowPyrLK(_np(prev), _np(nxt), p0, None, winSize=(ww, wh), maxLevel=max_level,
                                                   criteria=(ctype, ccount, ceps), flags=flags, minEigThreshold=min_eig)
        np.ctypeslib.as_array(next_xy, shape=(n * 2,))[:] = p1.reshape(-1)
        np.ctypeslib.as_array(status, shape=(n,))[:] = st.ravel()
        np.ctypeslib.as_array(err, shape=(n,))[:] = er.ravel()
        self._emit("pyr_lk", prev_pts=p0.copy(), next_pts=p1.reshape(n, 2).copy(), status=st.ravel().copy())

    def _estimate_affine_partial(self, from_xy, to_xy, n, method, thresh, max_iters, conf, refine, m6, mask, ok):
        a = np.ctypeslib.as_array(from_xy, shape=(n * 2,)).reshape(n, 2).astype(f32)
        b = np.ctypeslib.as_array(to_xy, shape=(n * 2,)).reshape(n, 2).astype(f32)
        M, inl = self.cv2.estimateAffinePartial2D(a, b, None, method, thresh, max_iters, conf, refine)
        if M is None or M.shape != (2, 3):
            ok[0] = 0
            self._emit("affine", affine=None, mask=None)
            return
        ok[0] = 1
        for i in range(6):
            m6[i] = float(M.ravel()[i])
        np.ctypeslib.as_array(mask, shape=(n,))[:] = inl.ravel()
        self._emit("affine", affine=M.copy(), mask=inl.ravel().copy())

    def _warp_affine(self, src, dst, m6, m_is_f32, flags, border_mode, border_value):
        M = np.array([m6[i] for i in range(6)], np.float64).reshape(2, 3)
        if m_is_f32:
            M = M.astype(f32)
        d = _np(dst)
        bv = tuple(border_value[i] for i in range(4))
        _put(dst, self.cv2.warpAffine(_np(src), M, (d.shape[1], d.shape[0]), dst=_dst(dst), flags=flags, borderMode=border_mode, borderValue=bv))
        self._emit("warp", T=M.copy())

    def _copy_make_border(self, src, dst, top, bottom, left, right, btype, value):
        bv = tuple(value[i] for i in range(4))
        _put(dst, self.cv2.copyMakeBorder(_np(src), top, bottom, left, right, btype, dst=_dst(dst), value=bv))

    def _add_weighted(self, a, alpha, b, beta, gamma, dst):
        # the Python restatement documents that cv::addWeighted is taken with optimisations ON (its SIMD path fuses
        # src1*alpha + src2*beta; plain path differs by <= 1 LSB) — a stock OpenCV build runs optimised
        self.cv2.setUseOptimized(True)
        _put(dst, self.cv2.addWeighted(_np(a), alpha, _np(b), beta, gamma))

    def _threshold(self, src, dst, thresh, maxval, ttype):
        _put(dst, self.cv2.threshold(_np(src), thresh, maxval, ttype)[1])

    def _find_contours(self, img, mode, method, pts, pts_cap, lens, lens_cap, ncont):
        cs, _ = self.cv2.findContours(np.ascontiguousarray(_np(img)), mode, method)
        k = 0
        if len(cs) > lens_cap:
            raise ValueError("findContours: too many contours")
        for i, c in enumerate(cs):
            c = c.reshape(-1, 2)
            if k + len(c) > pts_cap:
                raise ValueError("findContours: too many points")
            lens[i] = len(c)
            for p in c:
                pts[2 * k] = int(p[0])
                pts[2 * k + 1] = int(p[1])
                k += 1
        ncont[0] = len(cs)

    # ---- cv::KalmanFilter: state9 = statePre, statePost, transitionMatrix, measurementMatrix, processNoiseCov,
    #      measurementNoiseCov, errorCovPre, gain, errorCovPost
    _KF = ["statePre", "statePost", "transitionMatrix", "measurementMatrix", "processNoiseCov", "measurementNoiseCov",
           "errorCovPre", "gain", "errorCovPost"]

    def _kalman_create(self, dp, mp, cp, handle):
        h = self._next_handle
        self._next_handle += 1
        self._kalman[h] = self.cv2.KalmanFilter(dp, mp, cp)
        handle[0] = h

    def _kf_sync_in(self, h, state9):
        kf = self._kalman[h]
        for i, name in enumerate(self._KF):
            setattr(kf, name, np.ascontiguousarray(_np(state9[i])).copy())
        return kf

    def _kf_sync_out(self, kf, state9):
        for i, name in enumerate(self._KF):
            _put(state9[i], getattr(kf, name))

    def _kalman_predict(self, h, state9):
        kf = self._kf_sync_in(h, state9)
        kf.predict()
        self._kf_sync_out(kf, state9)

    def _kalman_correct(self, h, meas, state9):
        kf = self._kf_sync_in(h, state9)
        kf.correct(np.ascontiguousarray(_np(meas)).copy())
        self._kf_sync_out(kf, state9)

    def _kalman_release(self, h):
        self._kalman.pop(h, None)

    # ---- RollCorrection / AutoZoomCrop operations (cv::cuda:: in the reference -> the CPU functions of the same library)
    def _canny(self, src, dst, low, high, aperture, l2):
        g = np.ascontiguousarray(_np(src))
        _put(dst, self.cv2.Canny(g, low, high, apertureSize=aperture, L2gradient=bool(l2)))
        self._emit("canny", gray=g.copy(), edges=_np(dst).copy())

    def _hough_lines(self, edges, rho, theta, threshold, max_lines, out, votes, cap, n):
        lv = self.cv2.HoughLinesWithAccumulator(np.ascontiguousarray(_np(edges)), rho, theta, threshold)
        lv = np.zeros((0, 3), f32) if lv is None else lv.reshape(-1, 3).astype(f32)
        if max_lines > 0:
            lv = lv[:max_lines]
        lv = lv[:cap]
        n[0] = len(lv)
        if len(lv):
            np.ctypeslib.as_array(out, shape=(len(lv) * 2,))[:] = np.ascontiguousarray(lv[:, :2]).ravel()
            np.ctypeslib.as_array(votes, shape=(len(lv),))[:] = lv[:, 2].astype(np.int32)
        self._emit("hough", lines=lv[:, :2].copy(), votes=lv[:, 2].astype(np.int32), edges=_np(edges).copy())

    def _gaussian_blur(self, src, dst, kw, kh, sx, sy):
        _put(dst, self.cv2.GaussianBlur(_np(src), (kw, kh), sx, sigmaY=sy))

    def _remap(self, src, dst, mapx, mapy, interp, border):
        _put(dst, self.cv2.remap(_np(src), _np(mapx), _np(mapy), interp, borderMode=border))

    def _morphology(self, src, dst, op, kernel):
        _put(dst, self.cv2.morphologyEx(_np(src), op, np.ascontiguousarray(_np(kernel))))

    def _structuring_element(self, shape, dst):
        d = _np(dst)
        _put(dst, self.cv2.getStructuringElement(shape, (d.shape[1], d.shape[0])))

    def _draw_contours(self, img, pts, lens, ncont, idx, color, thickness):
        cs, k = [], 0
        for i in range(ncont):
            c = np.array([[pts[2 * (k + j)], pts[2 * (k + j) + 1]] for j in range(lens[i])], np.int32).reshape(-1, 1, 2)
            cs.append(c)
            k += lens[i]
        im = _np(img)
        self.cv2.drawContours(im, cs, idx, tuple(color[i] for i in range(4)), thickness)

    def _rotation_matrix(self, cx, cy, angle, scale, m6):
        M = self.cv2.getRotationMatrix2D((cx, cy), angle, scale)
        for i in range(6):
            m6[i] = float(M.ravel()[i])

    def _sobel(self, src, dst, dx, dy, ksize):
        d = _np(dst)
        depth = {np.dtype(np.int16): self.cv2.CV_16S, np.dtype(np.float32): self.cv2.CV_32F, np.dtype(np.uint8): self.cv2.CV_8U}[d.dtype]
        _put(dst, self.cv2.Sobel(_np(src), depth, dx, dy, ksize=ksize))


# ---------------------------------------------------------------------------------------------- library loading
def _declare(lib):
    vp, ci, cf, cz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    sig = {
        "mini_cv_set_ops": (None, [C.POINTER(OpsTable)]),
        "vsref_params_new": (vp, []), "vsref_params_delete": (None, [vp]),
        "vsref_params_set_num": (ci, [vp, C.c_char_p, C.c_double]),
        "vsref_params_get_num": (ci, [vp, C.c_char_p, PD]),
        "vsref_params_set_str": (ci, [vp, C.c_char_p, C.c_char_p]),
        "vsref_new": (vp, [vp]), "vsref_delete": (None, [vp]), "vsref_clean": (None, [vp]),
        "vsref_stabilize": (ci, [vp, vp, ci, ci, cz, vp, cz, PI, PI]),
        "vsref_flush": (ci, [vp, vp, cz, PI, PI]),
        "vsref_stabilize_nocopy": (ci, [vp, vp, ci, ci, cz, PI, PI]),
        "vsref_n_transforms": (ci, [vp]), "vsref_queue_size": (ci, [vp]), "vsref_smoothing_radius": (ci, [vp]),
        "vsref_get_transforms": (None, [vp, vp, ci]), "vsref_get_path": (None, [vp, vp, ci]),
        "vsref_n_smoothed": (ci, [vp]), "vsref_get_smoothed": (None, [vp, vp, ci]),
        "vsref_n_keypoints": (ci, [vp]), "vsref_get_keypoints": (None, [vp, vp, ci]),
        "vsref_set_transforms": (None, [vp, vp, ci]),
        "vsref_vc_last": (ci, [vp, vp, vp]),
        "vsref_vc_apply": (ci, [vp, vp, ci, ci, cz, vp, vp, cz, PI, PI]),
        "vsref_box_filter": (ci, [vp, vp, ci, vp]),
        "vsref_gaussian_filter": (ci, [vp, vp, ci, cf, vp]),
        "vsref_kalman_filter": (ci, [vp, vp, ci, vp]),
        "vsref_adaptive_radius": (ci, [vp, vp, vp, vp, ci]),
        "vsref_motion_intent": (ci, [vp, vp, ci]),
        "vsref_stabilization_strength": (cf, [vp, ci, vp]),
        "vsref_variance": (cf, [vp, vp, ci]), "vsref_consistency": (cf, [vp, vp, ci]),
        "vsref_adapt_smoothing_radius": (None, [vp, vp]),
        "vsref_drone_chain": (None, [vp, vp, vp]),
        "vsref_drone_analysis_size": (None, [vp, ci, ci, PI, PI]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def load_private_copy(path: str = LIB):
    """dlopen a private copy of the library: fresh function-local statics for this instance."""
    if not os.path.exists(path):
        raise FileNotFoundError(path + " — run `python oracle/build_ref.py` where /root/reference exists")
    fd, tmp = tempfile.mkstemp(prefix="vsref_", suffix=".so")
    os.close(fd)
    shutil.copyfile(path, tmp)
    try:
        lib = C.CDLL(tmp)
    finally:
        os.unlink(tmp)      # the mapping stays valid
    return lib


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class RefStabilizer:
    """`vs::Stabilizer` of the reference, run for real.  `params` is any object / dict with the reference's
    `Parameters` field names (e.g. `oracle.stabilizer_ref.Parameters`)."""

    def __init__(self, params=None, use_optimized: bool = False, record: bool = True, **kw):
        self.lib = _declare(load_private_copy())
        self.ops = CvOps(use_optimized)
        self.lib.mini_cv_set_ops(C.byref(self.ops.table))
        fields = {}
        if params is not None:
            fields.update(params if isinstance(params, dict) else vars(params))
        fields.update(kw)
        p = self.lib.vsref_params_new()
        for k, v in fields.items():
            if isinstance(v, str):
                rc = self.lib.vsref_params_set_str(p, k.encode(), v.encode())
            else:
                rc = self.lib.vsref_params_set_num(p, k.encode(), float(v))
            if rc != 0:
                raise KeyError(f"vs::Stabilizer::Parameters has no field {k!r}")
        self.h = self.lib.vsref_new(p)
        self.lib.vsref_params_delete(p)
        if not self.h:
            raise RuntimeError("reference Stabilizer constructor failed")
        self._out = None
        self._method = str(fields.get("smoothingMethod", "box"))
        self._n_out = 0
        # per generateTransform() call / per emitted frame, from the callbacks and the private members; the same
        # record types as the Python restatement so tests treat both alike
        self.frame_records: list[FrameRecord] = []
        self.output_records: list[OutputRecord] = []
        self.first_corners = None
        self._cur = None
        self._warp = None
        self.record = record
        if record:
            self.ops.log = self._on_op

    def _on_op(self, op, kw):
        if op == "warp":
            if self._warp is None:                 # the frame's own warp; the virtual canvas warps temporal frames after it
                self._warp = kw["T"]
            return
        if self._cur is None:
            return
        if op == "gftt":
            self._cur.setdefault("gftt", []).append(kw["corners"])
        elif op == "pyr_lk":
            self._cur.update(prev_pts=kw["prev_pts"], next_pts=kw["next_pts"], status=kw["status"])
        elif op == "affine":
            self._cur.update(affine=kw["affine"], inlier_mask=kw["mask"])

    def _buf(self, w, h, b=0):
        need = (w + 2 * b) * (h + 2 * b) * 3
        if self._out is None or self._out.size < need:
            self._out = np.empty(need, np.uint8)
        return self._out

    def _after_emit(self, produced: bool):
        """One OutputRecord per popped frame (applyNextSmoothTransform, Stabilizer.cpp:763-1137): smoothedPath_[idx] is a
        member; radius and intent are locals there, recomputed by calling the reference's own calculateAdaptiveRadius /
        analyzeMotionIntent on the same members; T is what reached cv::warpAffine."""
        if not produced or not self.record:
            return
        idx = self._n_out
        self._n_out += 1
        n = self.lib.vsref_n_transforms(self.h)
        if idx >= n or self._warp is None:                       # bounds-guard passthrough (:774-780)
            self.output_records.append(OutputRecord(idx, n, 0, 0, np.zeros(3, f32), None))
            self._warp = None
            return
        path = self.path()
        radius = 0
        sm = self.smoothed()[idx].copy()
        if self._method not in ("gaussian", "kalman"):
            radius = self.adaptive_radius(path[:, 0], path[:, 1], path[:, 2])
        intent = self.motion_intent(self.transforms()[idx], idx) if idx > 0 else 0
        self.output_records.append(OutputRecord(idx, n, radius, intent, sm, np.asarray(self._warp, f32)))
        self._warp = None

    def stabilize(self, frame: np.ndarray):
        if frame is None or frame.size == 0:
            self.lib.vsref_stabilize(self.h, None, 0, 0, 0, None, 0, None, None)
            return None
        frame = np.ascontiguousarray(frame)
        h, w = frame.shape[:2]
        first = self.lib.vsref_n_transforms(self.h) == 0 and self.lib.vsref_queue_size(self.h) == 0
        n0 = self.lib.vsref_n_transforms(self.h)
        self._cur = {}
        self._warp = None
        out = self._buf(w, h, 512)
        ow, oh = C.c_int(), C.c_int()
        rc = self.lib.vsref_stabilize(self.h, _fp(frame), w, h, frame.strides[0], _fp(out), out.size, C.byref(ow), C.byref(oh))
        cur, self._cur = self._cur, None
        if rc < 0:
            raise RuntimeError(f"reference stabilize failed rc={rc}: {self.ops.errors[-3:]}")
        if self.lib.vsref_n_transforms(self.h) > n0:            # a generateTransform() call happened
            g = cur.get("gftt", [])
            tr = self.transforms()
            self.frame_records.append(FrameRecord(
                len(tr), cur.get("prev_pts", np.zeros((0, 2), f32)), cur.get("next_pts", np.zeros((0, 2), f32)),
                cur.get("status", np.zeros((0,), np.uint8)), cur.get("inlier_mask"), cur.get("affine"),
                tr[-1].copy(), self.path()[-1].copy(), g[-1] if g else None))
        elif first and cur.get("gftt"):
            self.first_corners = cur["gftt"][-1]
        self._after_emit(rc == 1)
        if rc == 0:
            return None
        return out[: ow.value * oh.value * 3].reshape(oh.value, ow.value, 3).copy()

    def stabilize_nocopy(self, frame: np.ndarray) -> bool:
        """Timing path: no input copy (the caller keeps `frame` alive and unchanged while it is queued), no output copy,
        no records.  Returns whether a stabilized frame was produced."""
        ow, oh = C.c_int(), C.c_int()
        rc = self.lib.vsref_stabilize_nocopy(self.h, _fp(frame), frame.shape[1], frame.shape[0], frame.strides[0], C.byref(ow), C.byref(oh))
        if rc < 0:
            raise RuntimeError(f"reference stabilize failed rc={rc}: {self.ops.errors[-3:]}")
        return rc == 1

    def flush(self):
        if self.lib.vsref_queue_size(self.h) == 0:
            return None
        out = self._out
        ow, oh = C.c_int(), C.c_int()
        self._warp = None
        rc = self.lib.vsref_flush(self.h, _fp(out), out.size, C.byref(ow), C.byref(oh))
        if rc < 0:
            raise RuntimeError(f"reference flush failed rc={rc}: {self.ops.errors[-3:]}")
        self._after_emit(rc == 1)
        if rc == 0:
            return None
        return out[: ow.value * oh.value * 3].reshape(oh.value, ow.value, 3).copy()

    def clean(self):
        self.lib.vsref_clean(self.h)

    def _vec3(self, n, getter):
        a = np.zeros((n, 3), f32)
        if n:
            getter(self.h, _fp(a), n)
        return a

    def transforms(self):
        return self._vec3(self.lib.vsref_n_transforms(self.h), self.lib.vsref_get_transforms)

    def path(self):
        return self._vec3(self.lib.vsref_n_transforms(self.h), self.lib.vsref_get_path)

    def smoothed(self):
        return self._vec3(self.lib.vsref_n_smoothed(self.h), self.lib.vsref_get_smoothed)

    def smoothing_radius(self):
        return self.lib.vsref_smoothing_radius(self.h)

    def vc_apply(self, frame: np.ndarray, t3) -> np.ndarray:
        """The reference's virtual canvas stage on one frame with correction (dx, dy, da) (Stabilizer.cpp:1130-1134)."""
        frame = np.ascontiguousarray(frame)
        h, w = frame.shape[:2]
        t = np.ascontiguousarray(np.asarray(t3, f32).reshape(3))
        out = self._buf(w, h, 0)
        ow, oh = C.c_int(), C.c_int()
        rc = self.lib.vsref_vc_apply(self.h, _fp(frame), w, h, frame.strides[0], _fp(t), _fp(out), out.size, C.byref(ow), C.byref(oh))
        if rc != 1:
            raise RuntimeError(f"reference virtual canvas failed rc={rc}: {self.ops.errors[-3:]}")
        return out[: ow.value * oh.value * 3].reshape(oh.value, ow.value, 3).copy()

    def vc_last(self):
        """virtual canvas: (correction (dx, dy, da) of the frame emitted last, canvas scale, temporal buffer length)"""
        t = np.zeros(3, f32)
        sc = C.c_float()
        n = self.lib.vsref_vc_last(self.h, _fp(t), C.byref(sc))
        return t, sc.value, n

    # ---- the pure-host functions on their own
    def set_transforms(self, t):
        t = np.ascontiguousarray(t, f32)
        self.lib.vsref_set_transforms(self.h, _fp(t), len(t))

    def _filt(self, fn, path, *extra):
        path = np.ascontiguousarray(path, f32)
        out = np.zeros(len(path), f32)
        n = fn(self.h, _fp(path), len(path), *extra, _fp(out))
        if n < 0:
            raise RuntimeError("reference filter failed")
        return out[:n]

    def box_filter(self, path):
        return self._filt(self.lib.vsref_box_filter, path)

    def gaussian_filter(self, path, sigma):
        return self._filt(self.lib.vsref_gaussian_filter, path, C.c_float(sigma))

    def kalman_filter(self, path):
        return self._filt(self.lib.vsref_kalman_filter, path)

    def adaptive_radius(self, px, py, pa):
        px, py, pa = (np.ascontiguousarray(v, f32) for v in (px, py, pa))
        return self.lib.vsref_adaptive_radius(self.h, _fp(px), _fp(py), _fp(pa), len(px))

    def motion_intent(self, motion, frame_index):
        m = np.ascontiguousarray(motion, f32)
        return self.lib.vsref_motion_intent(self.h, _fp(m), int(frame_index))

    def stabilization_strength(self, intent, motion):
        m = np.ascontiguousarray(motion, f32)
        return f32(self.lib.vsref_stabilization_strength(self.h, int(intent), _fp(m)))

    def variance(self, v):
        v = np.ascontiguousarray(v, f32)
        return f32(self.lib.vsref_variance(self.h, _fp(v), len(v)))

    def consistency(self, v):
        v = np.ascontiguousarray(v, f32)
        return f32(self.lib.vsref_consistency(self.h, _fp(v), len(v)))

    def adapt_smoothing_radius(self, motion):
        m = np.ascontiguousarray(motion, f32)
        self.lib.vsref_adapt_smoothing_radius(self.h, _fp(m))
        return self.smoothing_radius()

    def drone_chain(self, t):
        a = np.ascontiguousarray(t, f32)
        out = np.zeros(3, f32)
        self.lib.vsref_drone_chain(self.h, _fp(a), _fp(out))
        return out

    def drone_analysis_size(self, w, h):
        aw, ah = C.c_int(), C.c_int()
        self.lib.vsref_drone_analysis_size(self.h, w, h, C.byref(aw), C.byref(ah))
        return aw.value, ah.value

    def close(self):
        if self.h:
            self.lib.vsref_delete(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_clip(frames, params, flush: bool = True, use_optimized: bool = False):
    """Same contract as oracle.stabilizer_ref.run_clip, on the real reference."""
    st = RefStabilizer(params, use_optimized=use_optimized)
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

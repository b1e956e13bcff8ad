"""Python mirror of the reference's `vs::Stabilizer` (include/video/Stabilizer.h:70-198) over the
C-ABI: same constructor parameters (field names of `Stabilizer::Parameters`), same methods
(`stabilize(frame) -> frame | None`, `flush()`, `clean()`), same "empty Mat" convention (None).
All arithmetic runs in the CUDA library; this file only marshals buffers."""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

from . import _capi
from ._capi import VsFrameRecord, VsOutputRecord, VsParams, check, lib


@dataclasses.dataclass
class Parameters:
    """vs::Stabilizer::Parameters (Stabilizer.h:76-175).  Inert fields are accepted and ignored,
    exactly as the reference does."""
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
    motionPrediction: bool = True
    horizonLock: bool = False
    featureDetector: int = 0
    orbFeatures: int = 500
    fastThreshold: int = 10
    useROI: bool = False
    roi: tuple = (0, 0, 0, 0)
    adaptiveSmoothing: bool = False
    minSmoothingRadius: int = 5
    maxSmoothingRadius: int = 50
    outlierThreshold: float = 3.0
    intentionalMotionThreshold: float = 20.0
    stageOneRadius: int = 10
    stageTwoRadius: int = 25
    useTemporalFiltering: bool = False
    temporalWindowSize: int = 5
    fadeAlpha: float = 0.1
    fadeDuration: int = 30
    motionThresholdLow: float = 5.0
    motionThresholdHigh: float = 20.0
    borderScaleFactor: float = 2.0
    rollCompensation: bool = True
    rollCompensationFactor: float = 0.75
    deepStabilization: bool = False
    modelPath: str = ""
    jitterFrequency: int = 3
    separateTranslationRotation: bool = True
    useImuData: bool = False
    enableVirtualCanvas: bool = False
    canvasScaleFactor: float = 1.5
    temporalBufferSize: int = 30
    canvasBlendWeight: float = 0.7
    adaptiveCanvasSize: bool = True
    maxCanvasScale: float = 2.0
    minCanvasScale: float = 1.2
    preserveEdgeQuality: bool = True
    edgeBlendRadius: int = 20
    droneHighFreqMode: bool = False
    hfShakePx: float = 1.5
    hfAnalysisMaxWidth: int = 960
    hfRotLPAlpha: float = 0.2
    enableConditionalCLAHE: bool = True
    hfDeadZoneThreshold: float = 2.0
    hfFreezeDuration: int = 10
    hfMotionAccumulatorDecay: float = 0.9

    _MAP = {
        "useCuda": "use_cuda", "logging": "logging", "smoothingRadius": "smoothing_radius", "maxCorners": "max_corners",
        "qualityLevel": "quality_level", "minDistance": "min_distance", "blockSize": "block_size",
        "borderType": "border_type", "borderSize": "border_size", "cropNZoom": "crop_n_zoom",
        "smoothingMethod": "smoothing_method", "gaussianSigma": "gaussian_sigma", "motionPrediction": "motion_prediction",
        "horizonLock": "horizon_lock", "featureDetector": "feature_detector", "orbFeatures": "orb_features",
        "fastThreshold": "fast_threshold", "useROI": "use_roi", "adaptiveSmoothing": "adaptive_smoothing",
        "minSmoothingRadius": "min_smoothing_radius", "maxSmoothingRadius": "max_smoothing_radius",
        "outlierThreshold": "outlier_threshold", "intentionalMotionThreshold": "intentional_motion_threshold",
        "stageOneRadius": "stage_one_radius", "stageTwoRadius": "stage_two_radius",
        "useTemporalFiltering": "use_temporal_filtering", "temporalWindowSize": "temporal_window_size",
        "fadeAlpha": "fade_alpha", "fadeDuration": "fade_duration", "motionThresholdLow": "motion_threshold_low",
        "motionThresholdHigh": "motion_threshold_high", "borderScaleFactor": "border_scale_factor",
        "rollCompensation": "roll_compensation", "rollCompensationFactor": "roll_compensation_factor",
        "deepStabilization": "deep_stabilization", "modelPath": "model_path", "jitterFrequency": "jitter_frequency",
        "separateTranslationRotation": "separate_translation_rotation", "useImuData": "use_imu_data",
        "enableVirtualCanvas": "enable_virtual_canvas", "canvasScaleFactor": "canvas_scale_factor",
        "temporalBufferSize": "temporal_buffer_size", "canvasBlendWeight": "canvas_blend_weight",
        "adaptiveCanvasSize": "adaptive_canvas_size", "maxCanvasScale": "max_canvas_scale",
        "minCanvasScale": "min_canvas_scale", "preserveEdgeQuality": "preserve_edge_quality",
        "edgeBlendRadius": "edge_blend_radius", "droneHighFreqMode": "drone_high_freq_mode", "hfShakePx": "hf_shake_px",
        "hfAnalysisMaxWidth": "hf_analysis_max_width", "hfRotLPAlpha": "hf_rot_lp_alpha",
        "enableConditionalCLAHE": "enable_conditional_clahe", "hfDeadZoneThreshold": "hf_dead_zone_threshold",
        "hfFreezeDuration": "hf_freeze_duration", "hfMotionAccumulatorDecay": "hf_motion_accumulator_decay",
    }

    def to_c(self) -> VsParams:
        p = VsParams()
        check(lib.vs_params_default(C.byref(p)))
        for py, cn in self._MAP.items():
            v = getattr(self, py)
            if isinstance(v, str):
                v = v.encode()
            elif isinstance(v, bool):
                v = int(v)
            setattr(p, cn, v)
        p.roi_x, p.roi_y, p.roi_width, p.roi_height = (int(x) for x in self.roi)
        return p

    @classmethod
    def from_c(cls, p: VsParams) -> "Parameters":
        kw = {}
        for py, cn in cls._MAP.items():
            v = getattr(p, cn)
            f = cls.__dataclass_fields__[py]
            if isinstance(v, bytes):
                v = v.decode()
            elif f.type in ("bool", bool):
                v = bool(v)
            elif dict(VsParams._fields_)[cn] is C.c_float:
                v = float(f"{v:.7g}")          # float32 field -> shortest faithful decimal
            kw[py] = v
        kw["roi"] = (p.roi_x, p.roi_y, p.roi_width, p.roi_height)
        return cls(**kw)

    @classmethod
    def from_yaml(cls, path: str) -> "Parameters":
        """Reads the `stabilizer:` section of a reference config.yaml (keys of examples/vsg.cpp:1003-1114)."""
        p = VsParams()
        check(lib.vs_params_default(C.byref(p)))
        check(lib.vs_params_from_yaml(path.encode(), C.byref(p)))
        return cls.from_c(p)

    @classmethod
    def from_yaml_string(cls, text: str) -> "Parameters":
        p = VsParams()
        check(lib.vs_params_default(C.byref(p)))
        check(lib.vs_params_from_yaml_string(text.encode(), C.byref(p)))
        return cls.from_c(p)


Parameters.__annotations__.pop("_MAP", None)


def _addr(a: np.ndarray) -> int:
    return a.ctypes.data


class Stabilizer:
    """vs::Stabilizer.  `stabilize(frame)` takes an HxWx3 uint8 BGR numpy array and returns the
    stabilized frame or None while the reference would return an empty cv::Mat."""

    def __init__(self, params: Parameters | None = None, device: int = 0):
        self.params = params or Parameters()
        self._h = C.c_void_p()
        cp = self.params.to_c()
        check(lib.vs_stabilizer_create(C.byref(cp), device, C.byref(self._h)))
        self._out = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.vs_stabilizer_destroy(h)
            self._h = None

    def _out_buf(self, w: int, h: int) -> np.ndarray:
        b = self.params.borderSize if (self.params.borderSize > 0 and not self.params.cropNZoom) else 0
        need = (h + 2 * b) * (w + 2 * b) * 3
        if self._out is None or self._out.size < need:
            self._out = np.empty(need, np.uint8)
        return self._out

    def stabilize(self, frame: np.ndarray | None):
        if frame is None or frame.size == 0:
            return None
        if frame.dtype != np.uint8 or frame.ndim != 3 or frame.shape[2] != 3:
            raise ValueError("frame must be HxWx3 uint8 (CV_8UC3 BGR)")
        if frame.strides[2] != 1 or frame.strides[1] != 3:
            frame = np.ascontiguousarray(frame)
        h, w = frame.shape[:2]
        out = self._out_buf(w, h)
        ow, oh, produced = C.c_int(), C.c_int(), C.c_int()
        check(lib.vs_stabilizer_push(self._h, _addr(frame), w, h, frame.strides[0], _addr(out), 0, out.size,
                                     C.byref(ow), C.byref(oh), C.byref(produced)))
        if not produced.value:
            return None
        return out[: oh.value * ow.value * 3].reshape(oh.value, ow.value, 3).copy()

    def flush(self):
        if self._out is None:
            return None
        out = self._out
        ow, oh, produced = C.c_int(), C.c_int(), C.c_int()
        check(lib.vs_stabilizer_flush(self._h, _addr(out), 0, out.size, C.byref(ow), C.byref(oh), C.byref(produced)))
        if not produced.value:
            return None
        return out[: oh.value * ow.value * 3].reshape(oh.value, ow.value, 3).copy()

    def clean(self):
        check(lib.vs_stabilizer_clean(self._h))

    # ---- pipelined host API: n stabilize() calls in one (copy-in / compute / copy-out overlap)
    def _border(self) -> int:
        p = self.params
        return p.borderSize if (p.borderSize > 0 and not p.cropNZoom) else 0

    def stabilize_many(self, frames: np.ndarray, out: np.ndarray | None = None):
        """frames: (n,H,W,3) uint8, C-contiguous (page-locked memory overlaps best).  Returns the produced frames as
        an (m,H',W',3) view of `out` (m <= n), exactly the non-empty results of n stabilize() calls in order."""
        if frames.dtype != np.uint8 or frames.ndim != 4 or frames.shape[3] != 3 or not frames.flags.c_contiguous:
            raise ValueError("frames must be a C-contiguous (n,H,W,3) uint8 array")
        n, h, w, _ = frames.shape
        b = self._border()
        cap = (h + 2 * b) * (w + 2 * b) * 3
        if out is None:
            out = np.empty((n, cap), np.uint8)
        if out.size < n * cap or not out.flags.c_contiguous:
            raise ValueError("out too small")
        self._many_shape = (h, w)
        ow, oh, m = C.c_int(), C.c_int(), C.c_int()
        check(lib.vs_stabilizer_push_many(self._h, _addr(frames), h * w * 3, n, w, h, w * 3, _addr(out), 0, cap,
                                          C.byref(ow), C.byref(oh), C.byref(m)))
        flat = out.reshape(-1)[: n * cap].reshape(n, cap)
        return flat[: m.value, : oh.value * ow.value * 3].reshape(m.value, oh.value, ow.value, 3) if m.value else flat[:0]

    def flush_many(self, max_frames: int = 64, out: np.ndarray | None = None):
        h, w = self._many_shape
        b = self._border()
        cap = (h + 2 * b) * (w + 2 * b) * 3
        if out is None:
            out = np.empty((max_frames, cap), np.uint8)
        ow, oh, m = C.c_int(), C.c_int(), C.c_int()
        check(lib.vs_stabilizer_flush_many(self._h, _addr(out), 0, cap, max_frames, C.byref(ow), C.byref(oh), C.byref(m)))
        flat = out.reshape(-1)[: max_frames * cap].reshape(max_frames, cap)
        return flat[: m.value, : oh.value * ow.value * 3].reshape(m.value, oh.value, ow.value, 3) if m.value else flat[:0]

    # ---- device-resident API (raw device pointers, e.g. torch tensor.data_ptr())
    def push_device(self, d_frame: int, w: int, h: int, stride: int, d_out: int, out_stride: int, out_capacity: int,
                    borrow: bool = False):
        ow, oh, produced = C.c_int(), C.c_int(), C.c_int()
        check(lib.vs_stabilizer_push_device(self._h, d_frame, w, h, stride, d_out, out_stride, out_capacity,
                                            1 if borrow else 0, C.byref(ow), C.byref(oh), C.byref(produced)))
        return (ow.value, oh.value) if produced.value else None

    def push_many_device(self, d_frames: int, frame_step: int, n: int, w: int, h: int, stride: int, d_out: int,
                         out_stride: int, out_frame_capacity: int, borrow: bool = False) -> int:
        """stabilize() over n device frames (frame k at d_frames + k * frame_step) in one call; returns the number of
        outputs written consecutively from d_out.  Asynchronous: sync() before reading them."""
        ow, oh, produced = C.c_int(), C.c_int(), C.c_int()
        check(lib.vs_stabilizer_push_many_device(self._h, d_frames, frame_step, n, w, h, stride, d_out, out_stride,
                                                 out_frame_capacity, 1 if borrow else 0, C.byref(ow), C.byref(oh), C.byref(produced)))
        return produced.value

    def flush_device(self, d_out: int, out_stride: int, out_capacity: int):
        ow, oh, produced = C.c_int(), C.c_int(), C.c_int()
        check(lib.vs_stabilizer_flush_device(self._h, d_out, out_stride, out_capacity,
                                             C.byref(ow), C.byref(oh), C.byref(produced)))
        return (ow.value, oh.value) if produced.value else None

    def sync(self):
        check(lib.vs_stabilizer_sync(self._h))

    def join(self):
        """Public stream waits for the handle's internal analysis / detection streams (no host block)."""
        check(lib.vs_stabilizer_join(self._h))

    @property
    def stream(self) -> int:
        return lib.vs_stabilizer_stream(self._h) or 0

    # ---- diagnostics used by the parity tests
    def counts(self):
        a, b = C.c_int(), C.c_int()
        check(lib.vs_stabilizer_counts(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def frame_record(self, i: int) -> VsFrameRecord:
        r = VsFrameRecord()
        check(lib.vs_stabilizer_frame_record(self._h, i, C.byref(r)))
        return r

    def output_record(self, i: int) -> VsOutputRecord:
        r = VsOutputRecord()
        check(lib.vs_stabilizer_output_record(self._h, i, C.byref(r)))
        return r

    def frame_points(self, i: int) -> dict:
        r = self.frame_record(i)
        n = max(r.n_prev_pts, 0)
        prev = np.zeros((n, 2), np.float32)
        nxt = np.zeros((n, 2), np.float32)
        status = np.zeros(n, np.uint8)
        mask = np.zeros(max(r.n_tracked, 0), np.uint8)
        det = np.zeros((max(r.n_detected, 0), 2), np.float32)
        check(lib.vs_stabilizer_frame_points(self._h, i, _addr(prev), _addr(nxt), _addr(status), _addr(mask), _addr(det)))
        return {"prev": prev, "next": nxt, "status": status,
                "inlier_mask": mask if r.n_inliers >= 0 else None,
                "detected": det if r.n_detected >= 0 else None}

    def first_corners(self) -> np.ndarray:
        n = C.c_int()
        buf = np.zeros((4096, 2), np.float32)
        check(lib.vs_stabilizer_first_corners(self._h, _addr(buf), 4096, C.byref(n)))
        return buf[: n.value].copy()

    def launch_count(self) -> int:
        n = C.c_uint64()
        check(lib.vs_stabilizer_launch_count(self._h, C.byref(n)))
        return n.value

    def set_timing(self, on: bool):
        check(lib.vs_stabilizer_set_timing(self._h, int(on)))

    def stage_times(self) -> dict:
        return _stage_times(None, lib.vs_stabilizer_stage_time, self._h)

    def wait_event(self, cuda_event: int) -> None:
        """Frames pushed after this call wait on the device for `cuda_event` (a cudaEvent_t handle, e.g.
        torch.cuda.Event().cuda_event): stream-ordered hand-off of frames produced on another stream."""
        check(lib.vs_stabilizer_wait_event(self._h, C.c_void_p(cuda_event)))

    def trace(self, capacity: int = 65536):
        """(n, 3) float32 array of (stage, start us, end us) for the stage launches timed since the last query."""
        import ctypes as C
        buf = (C.c_float * (3 * capacity))()
        n = C.c_int()
        check(lib.vs_stabilizer_trace(self._h, buf, capacity, C.byref(n)))
        return np.frombuffer(buf, np.float32, 3 * min(n.value, capacity)).reshape(-1, 3).copy()


STAGES = ("resize_gray", "pyrdown", "pyr_lk", "motion", "gftt", "warp", "copy_in", "copy_out")


def _stage_times(set_fn, get_fn, h):
    out = {}
    for i, name in enumerate(STAGES):
        ms, n = C.c_double(), C.c_longlong()
        check(get_fn(h, i, C.byref(ms), C.byref(n)))
        out[name] = {"ms": ms.value, "count": n.value}
    return out


class StabilizerBatch:
    """N independent streams advanced in lock-step on one GPU, one launch per stage for the whole
    batch (BASELINE config 4).  Frames and outputs are device pointers."""

    def __init__(self, params: Parameters | None, n_streams: int, device: int = 0):
        self.params = params or Parameters()
        self.n = n_streams
        self._h = C.c_void_p()
        cp = self.params.to_c()
        check(lib.vs_batch_create(C.byref(cp), device, n_streams, C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.vs_batch_destroy(h)
            self._h = None

    def push_device(self, d_frames, w, h, stride, d_outs, out_stride, out_capacity, borrow=False):
        fa = (C.c_void_p * self.n)(*d_frames)
        oa = (C.c_void_p * self.n)(*d_outs)
        ow, oh, produced = C.c_int(), C.c_int(), C.c_int()
        check(lib.vs_batch_push_device(self._h, fa, w, h, stride, oa, out_stride, out_capacity, 1 if borrow else 0,
                                       C.byref(ow), C.byref(oh), C.byref(produced)))
        return (ow.value, oh.value) if produced.value else None

    def flush_device(self, d_outs, out_stride, out_capacity):
        oa = (C.c_void_p * self.n)(*d_outs)
        ow, oh, produced = C.c_int(), C.c_int(), C.c_int()
        check(lib.vs_batch_flush_device(self._h, oa, out_stride, out_capacity, C.byref(ow), C.byref(oh), C.byref(produced)))
        return (ow.value, oh.value) if produced.value else None

    def build_pyramids(self, d_frames, w, h, stride):
        """Analysis-image build alone (gray + pyramid) for every stream: profiling / roofline."""
        fa = (C.c_void_p * self.n)(*d_frames)
        check(lib.vs_batch_build_pyramids(self._h, fa, w, h, stride))

    def build_levels(self, d_frames, w, h, stride, parts: int):
        """parts & 1: level 0 (resize + gray); parts & 2: both pyrDown levels."""
        fa = (C.c_void_p * self.n)(*d_frames)
        check(lib.vs_batch_build_levels(self._h, fa, w, h, stride, parts))

    def clip_analyze_device(self, d_frames, w, h, count, d_transforms_out):
        """Lock-step analysis of n temporal chunks of a clip (vs_batch_clip_analyze_device): d_frames[l] = device address
        of frame first_l - 2 of chunk l (first_l even, >= 4), d_transforms_out[l] = device address of count x 3 float32."""
        fa = (C.c_void_p * self.n)(*d_frames)
        oa = (C.c_void_p * self.n)(*d_transforms_out)
        check(lib.vs_batch_clip_analyze_device(self._h, fa, w, h, count, oa))

    def wait_event(self, cuda_event: int) -> None:
        check(lib.vs_batch_wait_event(self._h, C.c_void_p(cuda_event)))

    def sync(self):
        check(lib.vs_batch_sync(self._h))

    def join(self):
        check(lib.vs_batch_join(self._h))

    @property
    def stream(self) -> int:
        return lib.vs_batch_stream(self._h) or 0

    def launch_count(self) -> int:
        n = C.c_uint64()
        check(lib.vs_batch_launch_count(self._h, C.byref(n)))
        return n.value

    def set_timing(self, on: bool):
        check(lib.vs_batch_set_timing(self._h, int(on)))

    def stage_times(self) -> dict:
        return _stage_times(None, lib.vs_batch_stage_time, self._h)

    def counts(self, stream: int = 0):
        a, b = C.c_int(), C.c_int()
        check(lib.vs_batch_stream_counts(self._h, stream, C.byref(a), C.byref(b)))
        return a.value, b.value

    def frame_record(self, stream: int, i: int) -> VsFrameRecord:
        r = VsFrameRecord()
        check(lib.vs_batch_frame_record(self._h, stream, i, C.byref(r)))
        return r

    def output_record(self, stream: int, i: int) -> VsOutputRecord:
        r = VsOutputRecord()
        check(lib.vs_batch_output_record(self._h, stream, i, C.byref(r)))
        return r

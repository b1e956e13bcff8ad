"""ctypes binding of the C-ABI in include/vstab_b200.h (the binding a reference maintainer would
write for a Python host).  Loads video-stab_b200/libvstab_b200.so and fails
loudly when it is missing: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvstab_b200.so")


class VsRollParams(C.Structure):
    """vs_roll_params (include/vstab_b200.h) == vs::RollCorrection::Parameters"""
    _fields_ = [
        ("scale_factor", C.c_double), ("canny_threshold_low", C.c_double), ("canny_threshold_high", C.c_double),
        ("canny_aperture", C.c_int32), ("hough_rho", C.c_float), ("hough_theta", C.c_float), ("hough_threshold", C.c_int32),
        ("angle_filter_min", C.c_double), ("angle_filter_max", C.c_double), ("angle_smoothing_alpha", C.c_double),
        ("angle_decay", C.c_double), ("max_angle_change_deg", C.c_double),
    ]


class VsParams(C.Structure):
    _fields_ = [
        ("use_cuda", C.c_int32), ("logging", C.c_int32), ("smoothing_radius", C.c_int32), ("max_corners", C.c_int32),
        ("quality_level", C.c_double), ("min_distance", C.c_double), ("block_size", C.c_int32),
        ("border_type", C.c_char * 32), ("border_size", C.c_int32), ("crop_n_zoom", C.c_int32),
        ("smoothing_method", C.c_char * 32), ("gaussian_sigma", C.c_double), ("motion_prediction", C.c_int32),
        ("horizon_lock", C.c_int32), ("feature_detector", C.c_int32), ("orb_features", C.c_int32),
        ("fast_threshold", C.c_int32), ("use_roi", C.c_int32), ("roi_x", C.c_int32), ("roi_y", C.c_int32),
        ("roi_width", C.c_int32), ("roi_height", C.c_int32), ("adaptive_smoothing", C.c_int32),
        ("min_smoothing_radius", C.c_int32), ("max_smoothing_radius", C.c_int32), ("outlier_threshold", C.c_double),
        ("intentional_motion_threshold", C.c_double), ("stage_one_radius", C.c_int32), ("stage_two_radius", C.c_int32),
        ("use_temporal_filtering", C.c_int32), ("temporal_window_size", C.c_int32), ("fade_alpha", C.c_float),
        ("fade_duration", C.c_int32), ("motion_threshold_low", C.c_float), ("motion_threshold_high", C.c_float),
        ("border_scale_factor", C.c_float), ("roll_compensation", C.c_int32), ("roll_compensation_factor", C.c_double),
        ("deep_stabilization", C.c_int32), ("model_path", C.c_char * 256), ("jitter_frequency", C.c_int32),
        ("separate_translation_rotation", C.c_int32), ("use_imu_data", C.c_int32),
        ("enable_virtual_canvas", C.c_int32), ("canvas_scale_factor", C.c_float), ("temporal_buffer_size", C.c_int32),
        ("canvas_blend_weight", C.c_float), ("adaptive_canvas_size", C.c_int32), ("max_canvas_scale", C.c_float),
        ("min_canvas_scale", C.c_float), ("preserve_edge_quality", C.c_int32), ("edge_blend_radius", C.c_int32),
        ("drone_high_freq_mode", C.c_int32), ("hf_shake_px", C.c_float), ("hf_analysis_max_width", C.c_int32),
        ("hf_rot_lp_alpha", C.c_float), ("enable_conditional_clahe", C.c_int32), ("hf_dead_zone_threshold", C.c_float),
        ("hf_freeze_duration", C.c_int32), ("hf_motion_accumulator_decay", C.c_float),
    ]


class VsFrameRecord(C.Structure):
    _fields_ = [("frame_index", C.c_int32), ("n_prev_pts", C.c_int32), ("n_tracked", C.c_int32),
                ("n_inliers", C.c_int32), ("ransac_iters", C.c_int32), ("n_detected", C.c_int32),
                ("transform", C.c_float * 3), ("path", C.c_float * 3), ("affine", C.c_double * 6)]


class VsOutputRecord(C.Structure):
    _fields_ = [("index", C.c_int32), ("passthrough", C.c_int32), ("path_len", C.c_int32), ("radius", C.c_int32),
                ("intent", C.c_int32), ("smoothed", C.c_float * 3), ("T", C.c_float * 6)]


# every symbol include/vstab_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_U8P = C.c_void_p          # raw addresses (host numpy buffers or device pointers)
_I = C.c_int
_IP = C.POINTER(C.c_int)
_SZ = C.c_size_t
SYMBOLS = {
    "vs_params_default": (_I, [C.POINTER(VsParams)]),
    "vs_params_from_yaml": (_I, [C.c_char_p, C.POINTER(VsParams)]),
    "vs_params_from_yaml_string": (_I, [C.c_char_p, C.POINTER(VsParams)]),
    "vs_stabilizer_create": (_I, [C.POINTER(VsParams), _I, C.POINTER(_P)]),
    "vs_stabilizer_destroy": (None, [_P]),
    "vs_stabilizer_push": (_I, [_P, _U8P, _I, _I, _SZ, _U8P, _SZ, _SZ, _IP, _IP, _IP]),
    "vs_stabilizer_flush": (_I, [_P, _U8P, _SZ, _SZ, _IP, _IP, _IP]),
    "vs_stabilizer_clean": (_I, [_P]),
    "vs_stabilizer_push_device": (_I, [_P, _U8P, _I, _I, _SZ, _U8P, _SZ, _SZ, C.c_uint, _IP, _IP, _IP]),
    "vs_stabilizer_flush_device": (_I, [_P, _U8P, _SZ, _SZ, _IP, _IP, _IP]),
    "vs_stabilizer_push_many": (_I, [_P, _U8P, _SZ, _I, _I, _I, _SZ, _U8P, _SZ, _SZ, _IP, _IP, _IP]),
    "vs_stabilizer_push_many_device": (_I, [_P, _U8P, _SZ, _I, _I, _I, _SZ, _U8P, _SZ, _SZ, C.c_uint, _IP, _IP, _IP]),
    "vs_stabilizer_flush_many": (_I, [_P, _U8P, _SZ, _SZ, _I, _IP, _IP, _IP]),
    "vs_stabilizer_sync": (_I, [_P]),
    "vs_stabilizer_join": (_I, [_P]),
    "vs_stabilizer_stream": (_P, [_P]),
    "vs_stabilizer_counts": (_I, [_P, _IP, _IP]),
    "vs_stabilizer_frame_record": (_I, [_P, _I, C.POINTER(VsFrameRecord)]),
    "vs_stabilizer_output_record": (_I, [_P, _I, C.POINTER(VsOutputRecord)]),
    "vs_stabilizer_frame_points": (_I, [_P, _I, _P, _P, _P, _P, _P]),
    "vs_stabilizer_first_corners": (_I, [_P, _P, _I, _IP]),
    "vs_stabilizer_launch_count": (_I, [_P, C.POINTER(C.c_uint64)]),
    "vs_stabilizer_set_timing": (_I, [_P, _I]),
    "vs_stabilizer_stage_time": (_I, [_P, _I, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "vs_batch_set_timing": (_I, [_P, _I]),
    "vs_stabilizer_wait_event": (_I, [_P, _P]),
    "vs_stabilizer_trace": (_I, [_P, C.POINTER(C.c_float), _I, C.POINTER(C.c_int)]),
    "vs_batch_stage_time": (_I, [_P, _I, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "vs_clip_halo": (_I, [_I]),
    "vs_clip_analyze": (_I, [_P, _U8P, _I, _I, _I, _I, _P, _IP]),
    "vs_clip_render": (_I, [_P, _P, _I, _U8P, _I, _I, _I, _I, _U8P, _IP, _IP]),
    "vs_clip_analyze_device": (_I, [_P, _U8P, _I, _I, _I, _I, _P, _IP]),
    "vs_clip_set_transforms_device": (_I, [_P, _P, _I, _I, _I]),
    "vs_clip_render_prepared_device": (_I, [_P, _U8P, _I, _I, _I, _I, _U8P, _IP, _IP]),
    "vs_clip_render_device": (_I, [_P, _P, _I, _U8P, _I, _I, _I, _I, _U8P, _IP, _IP]),
    "vs_roll_params_default": (_I, [C.POINTER(VsRollParams)]),
    "vs_roll_params_from_yaml": (_I, [C.c_char_p, C.POINTER(VsRollParams)]),
    "vs_roll_params_from_yaml_string": (_I, [C.c_char_p, C.POINTER(VsRollParams)]),
    "vs_roll_create": (_I, [C.POINTER(VsRollParams), _I, C.POINTER(_P)]),
    "vs_roll_destroy": (None, [_P]),
    "vs_roll_correct": (_I, [_P, _U8P, _I, _I, _SZ, _U8P, _SZ]),
    "vs_roll_correct_device": (_I, [_P, _U8P, _I, _I, _SZ, _U8P, _SZ, _P]),
    "vs_roll_reset": (_I, [_P]),
    "vs_roll_state": (_I, [_P, C.POINTER(C.c_double), _IP, _IP, C.POINTER(C.c_uint64)]),
    "vs_roll_debug": (_I, [_P, _IP, _IP, _P, _P, _P, _I]),
    "vs_canvas_create": (_I, [C.POINTER(VsParams), _I, C.POINTER(_P)]),
    "vs_canvas_destroy": (None, [_P]),
    "vs_canvas_apply_device": (_I, [_P, _U8P, _I, _I, _SZ, C.POINTER(C.c_float), C.POINTER(C.c_float), _I, _U8P, _SZ, _P]),
    "vs_canvas_info": (_I, [_P, C.POINTER(C.c_float), _IP]),
    "vs_auto_zoom_crop": (_I, [_U8P, _I, _I, _SZ, C.c_double, _I, _U8P, _SZ, _SZ, _IP, _IP]),
    "vs_auto_zoom_crop_device": (_I, [_U8P, _I, _I, _SZ, C.c_double, _U8P, _SZ, _SZ, _IP, _IP, _P]),
    "vs_auto_zoom_rect_from_mask": (_I, [_U8P, _I, _I, _SZ, _IP, _IP, _IP, _IP, _IP]),
    "vs_k_find_external_contours": (_I, [_U8P, _I, _I, _SZ, _IP, _I, _IP, _I, _IP]),
    "vs_k_content_mask": (_I, [_U8P, _I, _I, _SZ, _U8P, _U8P, _P]),
    "vs_batch_create": (_I, [C.POINTER(VsParams), _I, _I, C.POINTER(_P)]),
    "vs_batch_destroy": (None, [_P]),
    "vs_batch_push_device": (_I, [_P, C.POINTER(_P), _I, _I, _SZ, C.POINTER(_P), _SZ, _SZ, C.c_uint, _IP, _IP, _IP]),
    "vs_batch_flush_device": (_I, [_P, C.POINTER(_P), _SZ, _SZ, _IP, _IP, _IP]),
    "vs_batch_build_pyramids": (_I, [_P, C.POINTER(_P), _I, _I, _SZ]),
    "vs_batch_build_levels": (_I, [_P, C.POINTER(_P), _I, _I, _SZ, _I]),
    "vs_batch_sync": (_I, [_P]),
    "vs_nv12_to_bgr_device": (_I, [_U8P, _SZ, _U8P, _SZ, _I, _I, _U8P, _SZ, _P]),
    "vs_bgr_to_nv12_device": (_I, [_U8P, _SZ, _I, _I, _U8P, _SZ, _U8P, _SZ, _P]),
    "vs_batch_join": (_I, [_P]),
    "vs_batch_wait_event": (_I, [_P, _P]),
    "vs_batch_clip_analyze_device": (_I, [_P, C.POINTER(_P), _I, _I, _I, C.POINTER(_P)]),
    "vs_batch_stream": (_P, [_P]),
    "vs_batch_launch_count": (_I, [_P, C.POINTER(C.c_uint64)]),
    "vs_batch_stream_counts": (_I, [_P, _I, _IP, _IP]),
    "vs_batch_frame_record": (_I, [_P, _I, _I, C.POINTER(VsFrameRecord)]),
    "vs_batch_output_record": (_I, [_P, _I, _I, C.POINTER(VsOutputRecord)]),
    "vs_k_warp_affine_bgr8": (_I, [_U8P, _I, _I, _SZ, _SZ, _U8P, _I, _I, _SZ, _SZ, _P, _I, _P]),
    "vs_k_gray_pyramid": (_I, [_U8P, _I, _I, _SZ, _I, _I, _U8P, _U8P, _U8P, _P]),
    "vs_k_resize_linear_u8": (_I, [_U8P, _I, _I, _SZ, _I, _U8P, _I, _I, _SZ, _P]),
    "vs_k_good_features": (_I, [_U8P, _I, _I, _I, C.c_double, C.c_double, _P, _I, _IP, _P]),
    "vs_k_good_features_block": (_I, [_U8P, _I, _I, _I, C.c_double, C.c_double, _I, _P, _I, _IP, _P]),
    "vs_k_pyr_lk": (_I, [_U8P, _U8P, _I, _I, _P, _I, _P, _P, _P]),
    "vs_k_estimate_affine_partial": (_I, [_P, _P, _I, _P, _P, _IP, _IP, _P]),
    "vs_k_warp_output": (_I, [_U8P, _I, _I, _SZ, _P, _I, _I, _I, _U8P, _SZ, _IP, _IP, _P]),
    "vs_last_error": (C.c_char_p, []),
    "vs_version": (C.c_char_p, []),
    "vs_abi_version": (_I, []),
}

STATUS_NAMES = {0: "VS_OK", 1: "VS_ERR_INVALID_ARG", 2: "VS_ERR_NO_DEVICE", 3: "VS_ERR_CUDA",
                4: "VS_ERR_OUT_OF_MEMORY", 5: "VS_ERR_BUFFER_TOO_SMALL", 6: "VS_ERR_IO", 7: "VS_ERR_UNSUPPORTED"}


class VsError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {msg}")
        self.status = status


def load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "video-stab_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()


def check(status: int) -> None:
    if status != 0:
        raise VsError(status, (lib.vs_last_error() or b"").decode())

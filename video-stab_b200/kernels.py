"""Single-kernel entry points of the C-ABI (vs_k_*), operating on torch CUDA tensors.  One per
OpenCV call the reference makes on the path; used by the per-kernel parity tests and the bench."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._capi import check, lib


def _t():
    import torch
    return torch


def warp_affine(src, T, out=None, stream: int = 0):
    """cv::warpAffine INTER_LINEAR/BORDER_CONSTANT.  src: (N,H,W,3) or (H,W,3) uint8 cuda tensor;
    T: (N,2,3) float32 (numpy/host)."""
    torch = _t()
    s = src if src.dim() == 4 else src.unsqueeze(0)
    n, h, w, _ = s.shape
    Tm = np.ascontiguousarray(np.asarray(T, np.float32).reshape(n, 6))
    if out is None:
        out = torch.empty_like(s)
    check(lib.vs_k_warp_affine_bgr8(s.data_ptr(), w, h, s.stride(1), s.stride(0), out.data_ptr(), w, h, out.stride(1),
                                    out.stride(0), Tm.ctypes.data, n, stream))
    return out if src.dim() == 4 else out[0]


def gray_pyramid(bgr, first_frame: bool = False):
    torch = _t()
    h, w, _ = bgr.shape
    aw, ah = (480, 270) if first_frame else (960, 540)
    l0 = torch.empty((ah, aw), dtype=torch.uint8, device=bgr.device)
    if first_frame:
        check(lib.vs_k_gray_pyramid(bgr.data_ptr(), w, h, bgr.stride(0), aw, ah, l0.data_ptr(), None, None, None))
        return [l0]
    l1 = torch.empty(((ah + 1) // 2, (aw + 1) // 2), dtype=torch.uint8, device=bgr.device)
    l2 = torch.empty(((l1.shape[0] + 1) // 2, (l1.shape[1] + 1) // 2), dtype=torch.uint8, device=bgr.device)
    check(lib.vs_k_gray_pyramid(bgr.data_ptr(), w, h, bgr.stride(0), aw, ah, l0.data_ptr(), l1.data_ptr(), l2.data_ptr(), None))
    return [l0, l1, l2]


def resize_linear(src, dsize):
    torch = _t()
    dw, dh = dsize
    ch = 1 if src.dim() == 2 else src.shape[2]
    sh, sw = src.shape[:2]
    shape = (dh, dw) if src.dim() == 2 else (dh, dw, ch)
    dst = torch.empty(shape, dtype=torch.uint8, device=src.device)
    torch.cuda.synchronize()
    check(lib.vs_k_resize_linear_u8(src.data_ptr(), sw, sh, src.stride(0), ch, dst.data_ptr(), dw, dh, dst.stride(0), None))
    torch.cuda.synchronize()
    return dst


def good_features(gray, max_corners: int, quality: float, min_dist: float, block_size: int = 3) -> np.ndarray:
    h, w = gray.shape
    cap = max_corners if max_corners > 0 else 2048
    out = np.zeros((cap, 2), np.float32)
    n = C.c_int()
    check(lib.vs_k_good_features_block(gray.data_ptr(), w, h, max_corners, quality, min_dist, block_size, out.ctypes.data, cap, C.byref(n), None))
    return out[: min(n.value, cap)].copy()


def pyr_lk(prev, nxt, pts: np.ndarray):
    h, w = prev.shape
    p = np.ascontiguousarray(np.asarray(pts, np.float32).reshape(-1, 2))
    n = len(p)
    out = np.zeros((n, 2), np.float32)
    st = np.zeros(n, np.uint8)
    check(lib.vs_k_pyr_lk(prev.data_ptr(), nxt.data_ptr(), w, h, p.ctypes.data, n, out.ctypes.data, st.ctypes.data, None))
    return out, st


def estimate_affine_partial(src: np.ndarray, dst: np.ndarray):
    a = np.ascontiguousarray(np.asarray(src, np.float32).reshape(-1, 2))
    b = np.ascontiguousarray(np.asarray(dst, np.float32).reshape(-1, 2))
    n = len(a)
    aff = np.zeros(6, np.float64)
    mask = np.zeros(max(n, 1), np.uint8)
    iters, ok = C.c_int(), C.c_int()
    check(lib.vs_k_estimate_affine_partial(a.ctypes.data, b.ctypes.data, n, aff.ctypes.data, mask.ctypes.data,
                                           C.byref(iters), C.byref(ok), None))
    if not ok.value:
        return None, mask[:n], iters.value
    return aff.reshape(2, 3), mask[:n], iters.value


def warp_output(src, T, mode: int, border_size: int = 0, border_mode: int = 0):
    """mode 0 plain / 1 copyMakeBorder+warp / 2 warp+crop+zoom (Stabilizer.cpp:981-990,1056-1060,1108-1124)."""
    torch = _t()
    h, w, _ = src.shape
    b = border_size if mode == 1 else 0
    dst = torch.empty((h + 2 * b, w + 2 * b, 3), dtype=torch.uint8, device=src.device)
    Tm = np.ascontiguousarray(np.asarray(T, np.float32).reshape(6))
    ow, oh = C.c_int(), C.c_int()
    torch.cuda.synchronize()
    check(lib.vs_k_warp_output(src.data_ptr(), w, h, src.stride(0), Tm.ctypes.data, mode, border_size, border_mode,
                               dst.data_ptr(), dst.stride(0), C.byref(ow), C.byref(oh), None))
    return dst


def nv12_to_bgr(nv12, out=None, stream: int = 0):
    """(H * 3 / 2, W) uint8 cuda tensor (Y plane, then the interleaved UV plane) -> (H, W, 3) BGR; cv2.COLOR_YUV2BGR_NV12."""
    torch = _t()
    h32, w = nv12.shape
    h = h32 * 2 // 3
    if out is None:
        out = torch.empty((h, w, 3), dtype=torch.uint8, device=nv12.device)
    check(lib.vs_nv12_to_bgr_device(nv12.data_ptr(), nv12.stride(0), nv12[h:].data_ptr(), nv12.stride(0), w, h,
                                    out.data_ptr(), out.stride(0), stream))
    return out


def bgr_to_nv12(bgr, out=None, stream: int = 0):
    """(H, W, 3) BGR uint8 cuda tensor -> (H * 3 / 2, W) NV12; cv2.COLOR_BGR2YUV_I420 with U and V interleaved."""
    torch = _t()
    h, w, _ = bgr.shape
    if out is None:
        out = torch.empty((h * 3 // 2, w), dtype=torch.uint8, device=bgr.device)
    check(lib.vs_bgr_to_nv12_device(bgr.data_ptr(), bgr.stride(0), w, h, out.data_ptr(), out.stride(0), out[h:].data_ptr(),
                                    out.stride(0), stream))
    return out

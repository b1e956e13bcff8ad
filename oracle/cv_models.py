"""numpy restatements of the seven OpenCV operations on the stabilization hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each function restates the
published OpenCV algorithm behind one reference call site in
`/root/reference/src/Stabilizer.cpp`; `tests/test_oracle_models.py` pins every one
of them bit-exactly against the real library (`cv2` 4.13.0, setUseOptimized(False)).
They are the *specification* the CUDA kernels in `video-stab_b200/csrc/` follow.

    bgr2gray            cv::cvtColor BGR2GRAY          Stabilizer.cpp:305,450
    resize_linear       cv::resize INTER_LINEAR        Stabilizer.cpp:304,449,602,1121
    pyr_down            cv::pyrDown (inside PyrLK)     Stabilizer.cpp:611
    lk_track            cv::calcOpticalFlowPyrLK       Stabilizer.cpp:611-619
    min_eigen_map/gftt  cv::goodFeaturesToTrack        Stabilizer.cpp:355-357,740-744
    estimate_affine_partial_2d  cv::estimateAffinePartial2D   Stabilizer.cpp:647-649
    warp_affine         cv::warpAffine INTER_LINEAR    Stabilizer.cpp:1056-1060
    copy_make_border    cv::copyMakeBorder             Stabilizer.cpp:982-987
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32
f64 = np.float64


# --------------------------------------------------------------------------- gray
def bgr2gray(bgr: np.ndarray) -> np.ndarray:
    """cv::cvtColor(BGR2GRAY) for 8-bit: 15-bit fixed point, round half up."""
    b = bgr[..., 0].astype(np.int32)
    g = bgr[..., 1].astype(np.int32)
    r = bgr[..., 2].astype(np.int32)
    return ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8)


# ------------------------------------------------------------------------- resize
def _linear_axis_table(src: int, dst: int):
    """Per-axis sample table of cv::resize INTER_LINEAR (8-bit fixed-point path).

    Returns (s0, s1, a0, a1): tap indices and int16 coefficients (scale 2048).
    `horizontal` semantics: the left tap index is clamped and the fraction zeroed at
    both ends (OpenCV's xofs/ialpha construction).
    """
    scale = 1.0 / (float(dst) / float(src))          # double, as cv::resize computes it
    d = np.arange(dst, dtype=np.float64)
    fx = ((d + 0.5) * scale - 0.5).astype(f32)       # (float)((dx+0.5)*scale_x - 0.5)
    s = np.floor(fx).astype(np.int32)
    fx = (fx - s.astype(f32)).astype(f32)
    return s, fx


def resize_linear(img: np.ndarray, dsize: tuple[int, int]) -> np.ndarray:
    """cv::resize(img, dsize=(W,H), INTER_LINEAR) for CV_8UC1 / CV_8UC3.

    * exact 2x2 decimation -> OpenCV switches to its INTER_AREA fast path:
      (p00+p01+p10+p11+2)>>2
    * everything else -> 11-bit fixed-point separable bilinear
      (HResizeLinear / VResizeLinear<uchar,int,short>).
    """
    dw, dh = dsize
    sh, sw = img.shape[:2]
    a = img if img.ndim == 3 else img[..., None]
    if sw == 2 * dw and sh == 2 * dh:
        p = a.astype(np.int32)
        out = (p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2] + 2) >> 2
        out = out.astype(np.uint8)
        return out if img.ndim == 3 else out[..., 0]

    # horizontal table: clamp index AND zero the fraction outside
    sx, fx = _linear_axis_table(sw, dw)
    lo = sx < 0
    sx = np.where(lo, 0, sx)
    fx = np.where(lo, f32(0), fx)
    hi = sx >= sw - 1
    sx = np.where(hi, sw - 1, sx)
    fx = np.where(hi, f32(0), fx)
    ax0 = np.rint((f32(1) - fx) * f32(2048)).astype(np.int32)
    ax1 = np.rint(fx * f32(2048)).astype(np.int32)
    sx1 = np.minimum(sx + 1, sw - 1)

    # vertical table: coefficients from the UNclamped fraction, row indices clipped
    sy, fy = _linear_axis_table(sh, dh)
    by0 = np.rint((f32(1) - fy) * f32(2048)).astype(np.int32)
    by1 = np.rint(fy * f32(2048)).astype(np.int32)
    sy0 = np.clip(sy, 0, sh - 1)
    sy1 = np.clip(sy + 1, 0, sh - 1)

    p = a.astype(np.int32)
    hrow = p[:, sx, :] * ax0[None, :, None] + p[:, sx1, :] * ax1[None, :, None]   # (sh, dw, c)
    h0 = hrow[sy0]
    h1 = hrow[sy1]
    out = (((by0[:, None, None] * (h0 >> 4)) >> 16) + ((by1[:, None, None] * (h1 >> 4)) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out if img.ndim == 3 else out[..., 0]


# ------------------------------------------------------------------------ pyrDown
def pyr_down(img: np.ndarray) -> np.ndarray:
    """cv::pyrDown 8UC1: separable [1 4 6 4 1], REFLECT_101, (sum+128)>>8, even samples."""
    h, w = img.shape
    p = np.pad(img.astype(np.int32), 2, mode="reflect")
    k = (1, 4, 6, 4, 1)
    r = sum(k[i] * p[:, i:i + w] for i in range(5))[:, ::2]
    c = sum(k[i] * r[i:i + h] for i in range(5))[::2]
    return ((c + 128) >> 8).astype(np.uint8)


def build_pyramid(img: np.ndarray, max_level: int = 2) -> list[np.ndarray]:
    pyr = [img]
    for _ in range(max_level):
        pyr.append(pyr_down(pyr[-1]))
    return pyr


# ----------------------------------------------------------------------------- LK
def scharr_deriv(img: np.ndarray):
    """cv::calcScharrDeriv: int16 unscaled Scharr, image borders REFLECT_101."""
    h, w = img.shape
    p = np.pad(img.astype(np.int32), 1, mode="reflect")

    def s(dy, dx):
        return p[1 + dy:1 + dy + h, 1 + dx:1 + dx + w]

    ix = 3 * (s(-1, 1) - s(-1, -1)) + 10 * (s(0, 1) - s(0, -1)) + 3 * (s(1, 1) - s(1, -1))
    iy = 3 * (s(1, -1) - s(-1, -1)) + 10 * (s(1, 0) - s(-1, 0)) + 3 * (s(1, 1) - s(-1, 1))
    return ix.astype(np.int16), iy.astype(np.int16)


_WIN = 15            # the reference always uses cv::Size(15,15)   (Stabilizer.cpp:616)
_HALF = f32(7.0)     # (win-1)/2
_W_BITS = 14
_FLT_SCALE = f32(1.0 / (1 << 20))


def _lk_weights(a, b):
    one = f32(1)
    s = f32(1 << _W_BITS)
    iw00 = int(np.rint((one - a) * (one - b) * s))
    iw01 = int(np.rint(a * (one - b) * s))
    iw10 = int(np.rint((one - a) * b * s))
    iw11 = (1 << _W_BITS) - iw00 - iw01 - iw10
    return iw00, iw01, iw10, iw11


def _lk_interp(plane, x0, y0, w4, shift):
    """(sum of 4 weighted taps + half) >> shift over the 15x15 window whose top-left
    integer corner is (x0,y0); `plane` is padded by _WIN on every side."""
    y0 += _WIN
    x0 += _WIN
    w = plane[y0:y0 + _WIN + 1, x0:x0 + _WIN + 1]
    v = w[:-1, :-1] * w4[0] + w[:-1, 1:] * w4[1] + w[1:, :-1] * w4[2] + w[1:, 1:] * w4[3]
    return (v + (1 << (shift - 1))) >> shift


def _acc_cov(u, v):
    """float32 accumulation order of OpenCV's 128-bit SIMD LK loop: columns 0..7 go to
    4 vector lanes (lane l sees x=l then x=l+4 of every row, product rounded to f32
    then added), columns 8..14 to one scalar chain; total = scalar + ((q0+q2)+(q1+q3))."""
    q = np.zeros(4, f32)
    s = f32(0)
    for y in range(_WIN):
        for h in (0, 4):
            fu = u[y, h:h + 4].astype(f32)
            fv = v[y, h:h + 4].astype(f32)
            q = (fu * fv).astype(f32) + q
        for x in range(8, _WIN):
            s = s + f32(int(u[y, x]) * int(v[y, x]))
    return s + ((q[0] + q[2]) + (q[1] + q[3]))


def _acc_b(diff, gx, gy):
    """float32 accumulation order of the mismatch vector (b1,b2): 8 vector chains of
    int32 pair sums (pixels (0,4),(1,5),(2,6),(3,7) of each row) + 2 scalar chains."""
    qb0 = np.zeros(4, f32)
    qb1 = np.zeros(4, f32)
    s1 = f32(0)
    s2 = f32(0)
    for y in range(_WIN):
        d = diff[y].astype(np.int64)
        x_ = gx[y].astype(np.int64)
        y_ = gy[y].astype(np.int64)
        qb0 = qb0 + np.array([d[0] * x_[0] + d[4] * x_[4], d[0] * y_[0] + d[4] * y_[4],
                              d[1] * x_[1] + d[5] * x_[5], d[1] * y_[1] + d[5] * y_[5]]).astype(f32)
        qb1 = qb1 + np.array([d[2] * x_[2] + d[6] * x_[6], d[2] * y_[2] + d[6] * y_[6],
                              d[3] * x_[3] + d[7] * x_[7], d[3] * y_[3] + d[7] * y_[7]]).astype(f32)
        for x in range(8, _WIN):
            s1 = s1 + f32(int(d[x] * x_[x]))
            s2 = s2 + f32(int(d[x] * y_[x]))
    qs = qb0 + qb1
    return s1 + (qs[0] + qs[2]), s2 + (qs[1] + qs[3])


def lk_track(prev: np.ndarray, nxt: np.ndarray, pts: np.ndarray, max_level: int = 2,
             max_count: int = 20, eps: float = 0.03, min_eig_thr: float = 1e-4,
             pyr_prev=None, pyr_next=None):
    """cv::calcOpticalFlowPyrLK(prev, next, pts, winSize=(15,15), maxLevel, criteria=
    (COUNT+EPS, max_count, eps), flags=0, minEigThreshold).  Returns (nextPts f32 Nx2,
    status u8 N).  Bit-exact restatement incl. the float accumulation order."""
    pyr_i = pyr_prev if pyr_prev is not None else build_pyramid(prev, max_level)
    pyr_j = pyr_next if pyr_next is not None else build_pyramid(nxt, max_level)
    pts = np.asarray(pts, f32).reshape(-1, 2)
    n = len(pts)
    status = np.ones(n, np.uint8)
    next_pts = np.zeros((n, 2), f32)
    eps2 = eps * eps
    for level in range(max_level, -1, -1):
        img_i = pyr_i[level]
        img_j = pyr_j[level]
        h, w = img_i.shape
        ip = np.pad(img_i, _WIN, mode="reflect").astype(np.int32)
        jp = np.pad(img_j, _WIN, mode="reflect").astype(np.int32)
        ix, iy = scharr_deriv(img_i)
        dxp = np.pad(ix.astype(np.int32), _WIN)       # derivative plane: zero outside
        dyp = np.pad(iy.astype(np.int32), _WIN)
        sc = f32(1.0 / (1 << level))
        for p in range(n):
            prev_pt = pts[p] * sc
            nxt_pt = prev_pt.copy() if level == max_level else next_pts[p] * f32(2)
            next_pts[p] = nxt_pt
            prev_pt = prev_pt - _HALF
            ipx = int(math.floor(prev_pt[0]))
            ipy = int(math.floor(prev_pt[1]))
            if ipx < -_WIN or ipx >= w or ipy < -_WIN or ipy >= h:
                if level == 0:
                    status[p] = 0
                continue
            a = f32(prev_pt[0] - f32(ipx))
            b = f32(prev_pt[1] - f32(ipy))
            w4 = _lk_weights(a, b)
            iw = _lk_interp(ip, ipx, ipy, w4, _W_BITS - 5)
            ixw = _lk_interp(dxp, ipx, ipy, w4, _W_BITS)
            iyw = _lk_interp(dyp, ipx, ipy, w4, _W_BITS)
            a11 = _acc_cov(ixw, ixw) * _FLT_SCALE
            a12 = _acc_cov(ixw, iyw) * _FLT_SCALE
            a22 = _acc_cov(iyw, iyw) * _FLT_SCALE
            det = a11 * a22 - a12 * a12
            min_eig = (a22 + a11 - np.sqrt((a11 - a22) * (a11 - a22) + f32(4) * a12 * a12)) / f32(2 * _WIN * _WIN)
            if float(min_eig) < min_eig_thr or det < np.finfo(f32).eps:
                if level == 0:
                    status[p] = 0
                continue
            det = f32(1) / det
            nxt_pt = nxt_pt - _HALF
            prev_delta = np.zeros(2, f32)
            for j in range(max_count):
                inx = int(math.floor(nxt_pt[0]))
                iny = int(math.floor(nxt_pt[1]))
                if inx < -_WIN or inx >= w or iny < -_WIN or iny >= h:
                    if level == 0:
                        status[p] = 0
                    break
                a = f32(nxt_pt[0] - f32(inx))
                b = f32(nxt_pt[1] - f32(iny))
                w4 = _lk_weights(a, b)
                jw = _lk_interp(jp, inx, iny, w4, _W_BITS - 5)
                ib1, ib2 = _acc_b(jw - iw, ixw, iyw)
                b1 = ib1 * _FLT_SCALE
                b2 = ib2 * _FLT_SCALE
                delta = np.array([(a12 * b2 - a22 * b1) * det, (a12 * b1 - a11 * b2) * det], f32)
                nxt_pt = nxt_pt + delta
                next_pts[p] = nxt_pt + _HALF
                if float(delta[0]) * float(delta[0]) + float(delta[1]) * float(delta[1]) <= eps2:
                    break
                if j > 0 and abs(delta[0] + prev_delta[0]) < 0.01 and abs(delta[1] + prev_delta[1]) < 0.01:
                    next_pts[p] = next_pts[p] - delta * f32(0.5)
                    break
                prev_delta = delta
    return next_pts, status


# --------------------------------------------------------------------------- GFTT
def sobel_scaled(img: np.ndarray, block_size: int = 3):
    """cv::Sobel(img, CV_32F, ksize=3, scale=1/(4*block*255)) in both directions with
    the float op order of OpenCV's baseline (non-FMA) separable filter engine."""
    s = f32(1.0 / (4.0 * block_size * 255.0))
    f0 = f32(2) * s
    f1 = s
    p = np.pad(img, 1, mode="reflect").astype(f32)
    cm, c0, cp = p[:, :-2], p[:, 1:-1], p[:, 2:]
    rx = cp - cm                                         # row pass of Dx: exact
    dx = (rx[:-2] + rx[2:]) * f1 + rx[1:-1] * f0         # column pass [1 2 1]*scale
    ry = (cm * f1 + c0 * f0) + cp * f1                   # row pass of Dy: [1 2 1]*scale, left-to-right
    dy = ry[2:] - ry[:-2]                                # column pass [-1 0 1]: exact
    return dx.astype(f32), dy.astype(f32)


def min_eigen_map(img: np.ndarray, block_size: int = 3) -> np.ndarray:
    """cv::cornerMinEigenVal(img, blockSize=3, ksize=3, BORDER_REFLECT_101)."""
    dx, dy = sobel_scaled(img, block_size)

    def box(a):                                          # un-normalised 3x3 box, double sums
        q = np.pad(a, 1, mode="reflect").astype(f64)
        r = q[:, :-2] + q[:, 1:-1] + q[:, 2:]
        return (r[:-2] + r[1:-1] + r[2:]).astype(f32)

    a = box(dx * dx) * f32(0.5)
    b = box(dx * dy)
    c = box(dy * dy) * f32(0.5)
    return ((a + c) - np.sqrt((a - c) * (a - c) + b * b)).astype(f32)


def gftt(img: np.ndarray, max_corners: int, quality: float, min_dist: float,
         block_size: int = 3, return_stats: bool = False):
    """cv::goodFeaturesToTrack(img, maxCorners, quality, minDistance, noArray, blockSize,
    useHarris=false).  Returns Nx2 float32 (x,y) in acceptance order."""
    if block_size != 3:
        raise NotImplementedError("the reference path only ever uses blockSize=3 (config default / :744)")
    e = min_eigen_map(img, block_size)
    h, w = e.shape
    thr = f32(f64(e.max()) * quality)
    e2 = np.where(e > thr, e, f32(0))
    p = np.pad(e2, 1, mode="constant", constant_values=-np.inf)
    dil = np.max([p[i:i + h, j:j + w] for i in range(3) for j in range(3)], axis=0)
    cand = (e2 != 0) & (e2 == dil)
    cand[0, :] = cand[-1, :] = False
    cand[:, 0] = cand[:, -1] = False
    ys, xs = np.nonzero(cand)
    vals = e2[ys, xs]
    addr = ys * w + xs
    order = np.lexsort((-addr, -vals.astype(f64)))       # value desc, then address desc
    ys, xs = ys[order], xs[order]
    out = []
    used = 0
    if min_dist >= 1:
        cell = int(np.rint(min_dist))
        gw = (w + cell - 1) // cell
        gh = (h + cell - 1) // cell
        grid = [[] for _ in range(gw * gh)]
        md2 = min_dist * min_dist
        for y, x in zip(ys.tolist(), xs.tolist()):
            used += 1
            cx, cy = x // cell, y // cell
            good = True
            for yy in range(max(0, cy - 1), min(gh - 1, cy + 1) + 1):
                for xx in range(max(0, cx - 1), min(gw - 1, cx + 1) + 1):
                    for (px, py) in grid[yy * gw + xx]:
                        if (x - px) ** 2 + (y - py) ** 2 < md2:
                            good = False
                            break
                    if not good:
                        break
                if not good:
                    break
            if good:
                grid[cy * gw + cx].append((x, y))
                out.append((x, y))
                if max_corners > 0 and len(out) == max_corners:
                    break
    else:
        for y, x in zip(ys.tolist(), xs.tolist()):
            used += 1
            out.append((x, y))
            if max_corners > 0 and len(out) == max_corners:
                break
    res = np.array(out, f32).reshape(-1, 2)
    if return_stats:
        return res, {"candidates": int(len(ys)), "visited": used}
    return res


# ------------------------------------------------------------------------- RANSAC
_RNG_COEFF = 4164903690
_DBL_MIN = 2.2250738585072014e-308


class CvRNG:
    """cv::RNG (multiply-with-carry)."""

    def __init__(self, state: int = 0xFFFFFFFFFFFFFFFF):
        self.state = state & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        self.state = ((self.state & 0xFFFFFFFF) * _RNG_COEFF + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a: int, b: int) -> int:
        return a if a == b else self.next() % (b - a) + a


def _ransac_update_num_iters(p: float, ep: float, model_points: int, max_iters: int) -> int:
    p = max(p, 0.0)
    p = min(p, 1.0)
    ep = max(ep, 0.0)
    ep = min(ep, 1.0)
    num = max(1.0 - p, _DBL_MIN)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < _DBL_MIN:
        return 0
    num = math.log(num)
    denom = math.log(denom)
    return max_iters if (denom >= 0 or -num >= max_iters * (-denom)) else int(np.rint(num / denom))


def partial_affine_from_2(src2, dst2):
    """AffinePartial2DEstimatorCallback::runKernel: exact 2-point similarity, double."""
    x1, y1 = float(src2[0][0]), float(src2[0][1])
    x2, y2 = float(src2[1][0]), float(src2[1][1])
    X1, Y1 = float(dst2[0][0]), float(dst2[0][1])
    X2, Y2 = float(dst2[1][0]), float(dst2[1][1])
    with np.errstate(all="ignore"):
        d = f64(1.0) / f64((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2))
        s0 = d * ((X1 - X2) * (x1 - x2) + (Y1 - Y2) * (y1 - y2))
        s1 = d * ((Y1 - Y2) * (x1 - x2) - (X1 - X2) * (y1 - y2))
        s2 = d * ((Y1 - Y2) * (x1 * y2 - x2 * y1) - (X1 * y2 - X2 * y1) * (y1 - y2) - (X1 * x2 - X2 * x1) * (x1 - x2))
        s3 = d * (-(X1 - X2) * (x1 * y2 - x2 * y1) - (Y1 * x2 - Y2 * x1) * (x1 - x2) - (Y1 * y2 - Y2 * y1) * (y1 - y2))
    return np.array([[s0, -s1, s2], [s1, s0, s3]], f64)


def affine_residuals(model, src, dst):
    """Affine2DEstimatorCallback::computeError — float32, no FMA."""
    m = model.astype(f32).ravel()
    a = (m[0] * src[:, 0] + m[1] * src[:, 1] + m[2]) - dst[:, 0]
    b = (m[3] * src[:, 0] + m[4] * src[:, 1] + m[5]) - dst[:, 1]
    return (a * a + b * b).astype(f32)


def similarity_lsq(src, dst):
    """Linear least-squares 4-DOF similarity on (src->dst) in double: the fixed point
    OpenCV's 10-iteration LM refinement converges to (problem is linear)."""
    x = src[:, 0].astype(f64)
    y = src[:, 1].astype(f64)
    u = dst[:, 0].astype(f64)
    v = dst[:, 1].astype(f64)
    mx, my, mu, mv = x.mean(), y.mean(), u.mean(), v.mean()
    xc, yc, uc, vc = x - mx, y - my, u - mu, v - mv
    den = (xc * xc + yc * yc).sum()
    a = (xc * uc + yc * vc).sum() / den
    b = (xc * vc - yc * uc).sum() / den
    tx = mu - (a * mx - b * my)
    ty = mv - (b * mx + a * my)
    return np.array([[a, -b, tx], [b, a, ty]], f64)


def estimate_affine_partial_2d(src, dst, thresh: float = 5.0, max_iters: int = 500,
                               confidence: float = 0.99, return_stats: bool = False):
    """cv::estimateAffinePartial2D(src, dst, noArray, RANSAC, 5.0, 500) (defaults:
    confidence 0.99, refineIters 10).  Returns (2x3 f64 or None, inlier mask u8)."""
    src = np.asarray(src, f32).reshape(-1, 2)
    dst = np.asarray(dst, f32).reshape(-1, 2)
    n = len(src)
    rng = CvRNG()
    best_mask = np.zeros(n, np.uint8)
    best_model = None
    max_good = 0
    niters = max_iters
    t = f32(thresh * thresh)
    it = 0
    iters_run = 0
    if n < 2:
        return (None, best_mask, {"iters": 0}) if return_stats else (None, best_mask)
    while it < niters:
        if n > 2:
            idx = []
            attempts = 0
            while len(idx) < 2 and attempts < 10000:       # getSubset (never degenerate for 2 pts)
                k = rng.uniform(0, n)
                if k in idx:
                    continue
                idx.append(k)
            ms1, ms2 = src[idx], dst[idx]
        else:
            ms1, ms2 = src, dst
        model = partial_affine_from_2(ms1, ms2)
        err = affine_residuals(model, src, dst)
        with np.errstate(invalid="ignore"):
            mask = (err <= t).astype(np.uint8)
        good = int(mask.sum())
        if good > max(max_good, 1):
            best_mask = mask
            best_model = model
            max_good = good
            niters = _ransac_update_num_iters(confidence, float(n - good) / n, 2, niters)
        it += 1
        iters_run += 1
    if best_model is None:
        return (None, best_mask, {"iters": iters_run}) if return_stats else (None, best_mask)
    inl = best_mask.astype(bool)
    refined = similarity_lsq(src[inl], dst[inl])
    if return_stats:
        return refined, best_mask, {"iters": iters_run, "ransac_model": best_model}
    return refined, best_mask


# --------------------------------------------------------------------------- warp
def invert_affine_f32(t: np.ndarray) -> np.ndarray:
    """cv::warpAffine's in-place inversion of the (float32 -> double) 2x3 matrix."""
    m = np.asarray(t, f32).astype(f64).ravel().copy()
    det = m[0] * m[4] - m[1] * m[3]
    det = 1.0 / det if det != 0.0 else 0.0
    a11 = m[4] * det
    a22 = m[0] * det
    m[0] = a11
    m[1] *= -det
    m[3] *= -det
    m[4] = a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2] = b1
    m[5] = b2
    return m


def warp_affine(src: np.ndarray, t: np.ndarray, dsize: tuple[int, int] | None = None) -> np.ndarray:
    """cv::warpAffine(src, T(2x3 f32), dsize, INTER_LINEAR, BORDER_CONSTANT(0)) 8UC3/8UC1."""
    sh, sw = src.shape[:2]
    dw, dh = dsize if dsize is not None else (sw, sh)
    m = invert_affine_f32(t)
    xs = np.arange(dw, dtype=f64)
    ys = np.arange(dh, dtype=f64)
    adelta = np.rint(m[0] * xs * 1024.0).astype(np.int64)
    bdelta = np.rint(m[3] * xs * 1024.0).astype(np.int64)
    x0 = np.rint((m[1] * ys + m[2]) * 1024.0).astype(np.int64) + 16
    y0 = np.rint((m[4] * ys + m[5]) * 1024.0).astype(np.int64) + 16
    X = (x0[:, None] + adelta[None, :]) >> 5
    Y = (y0[:, None] + bdelta[None, :]) >> 5
    sx = np.clip(X >> 5, -32768, 32767)
    sy = np.clip(Y >> 5, -32768, 32767)
    ax = X & 31
    ay = Y & 31
    a = src if src.ndim == 3 else src[..., None]
    pad = np.zeros((sh + 2, sw + 2, a.shape[2]), np.int64)
    pad[1:-1, 1:-1] = a
    # taps outside the source contribute 0: clip indices into the zero ring
    sxc0 = np.clip(sx + 1, 0, sw + 1)
    sxc1 = np.clip(sx + 2, 0, sw + 1)
    syc0 = np.clip(sy + 1, 0, sh + 1)
    syc1 = np.clip(sy + 2, 0, sh + 1)
    w00 = ((32 - ax) * (32 - ay) * 32)[..., None]
    w01 = (ax * (32 - ay) * 32)[..., None]
    w10 = ((32 - ax) * ay * 32)[..., None]
    w11 = (ax * ay * 32)[..., None]
    acc = pad[syc0, sxc0] * w00 + pad[syc0, sxc1] * w01 + pad[syc1, sxc0] * w10 + pad[syc1, sxc1] * w11
    out = ((acc + 16384) >> 15).astype(np.uint8)
    return out if src.ndim == 3 else out[..., 0]


# ------------------------------------------------------------------------- border
BORDER_CONSTANT, BORDER_REPLICATE, BORDER_REFLECT, BORDER_WRAP, BORDER_REFLECT_101 = 0, 1, 2, 3, 4


def border_interpolate(p: np.ndarray, length: int, mode: int) -> np.ndarray:
    """cv::borderInterpolate, vectorised; returns -1 for BORDER_CONSTANT outside."""
    p = np.asarray(p, np.int64).copy()
    if mode == BORDER_CONSTANT:
        return np.where((p >= 0) & (p < length), p, -1)
    if mode == BORDER_REPLICATE:
        return np.clip(p, 0, length - 1)
    if mode == BORDER_WRAP:
        return np.mod(p, length)
    if length == 1:
        return np.zeros_like(p)
    delta = 1 if mode == BORDER_REFLECT_101 else 0
    for _ in range(64):
        lo = p < 0
        p = np.where(lo, -p - 1 + delta, p)
        hi = p >= length
        p = np.where(hi, length - 1 - (p - length) - delta, p)
        if not ((p < 0) | (p >= length)).any():
            break
    return p


def copy_make_border(src: np.ndarray, b: int, mode: int) -> np.ndarray:
    """cv::copyMakeBorder(src, b,b,b,b, mode, Scalar(0,0,0))."""
    h, w = src.shape[:2]
    ys = border_interpolate(np.arange(-b, h + b), h, mode)
    xs = border_interpolate(np.arange(-b, w + b), w, mode)
    out = src[np.clip(ys, 0, h - 1)][:, np.clip(xs, 0, w - 1)].copy()
    if mode == BORDER_CONSTANT:
        out[ys < 0] = 0
        out[:, xs < 0] = 0
    return out


# ---------------------------------------------------------------------------------------------------------------
# libm / cv::addWeighted models used by the output stage (Stabilizer.cpp:902-906, 964-969)
def glibc_cosf_sinf(x):
    """glibc >= 2.28 cosf / sinf for |x| < pi/4 (sysdeps/ieee754/flt-32/s_sincosf.h, the ARM optimized-routines
    polynomial evaluated in double): returns (cos, sin) as float32.  Pinned bit-exactly against libm.so.6 in
    tests/test_oracle_models.py; the device code (csrc/k_motion.cu f_cos / f_sin) follows the same steps."""
    x = np.float32(x)
    ax = abs(float(x))
    assert ax < 0.7853981633974483
    if ax < 2.0 ** -12:
        return np.float32(1.0), x
    H = float.fromhex
    xd = np.float64(x)
    x2 = xd * xd
    c1, c2, c3, c4 = H('-0x1.ffffffd0c621cp-2'), H('0x1.55553e1068f19p-5'), H('-0x1.6c087e89a359dp-10'), H('0x1.99343027bf8c3p-16')
    s1, s2, s3 = H('-0x1.555545995a603p-3'), H('0x1.1107605230bc4p-7'), H('-0x1.994eb3774cf24p-13')
    x4 = x2 * x2
    C2 = c3 + x2 * c4
    C1 = c1 + x2 * c2
    x6 = x4 * x2
    c = 1.0 + x2 * C1
    cosv = np.float32(c + x6 * C2)
    x3 = xd * x2
    S1 = s2 + x2 * s3
    x7 = x3 * x2
    s = xd + x3 * s1
    sinv = np.float32(s + x7 * S1)
    return cosv, sinv


def add_weighted_u8(a, alpha, b, beta):
    """cv::addWeighted(a, alpha, b, beta, 0) on CV_8U with OpenCV's optimised (SIMD) path: the scalars are float32 and
    each element is rint(fma(a, alpha, b * beta)) saturated - the multiply-add is FUSED (the plain C++ path rounds
    a*alpha separately and can differ by 1 LSB).  Pinned against cv2 4.13.0 with setUseOptimized(True)."""
    al, be = np.float32(alpha), np.float32(beta)
    t = (a.astype(np.float64) * np.float64(al) + (b.astype(np.float32) * be).astype(np.float32).astype(np.float64)).astype(np.float32)
    return np.clip(np.rint(t), 0, 255).astype(np.uint8)


def glibc_atan2f_small(y, x):
    """glibc atan2f(y, x) for x > 0 and |y/x| < 7/16 (the rotation angle of a partial-affine fit): atanf of the
    quotient ROUNDED TO FLOAT.  The rounding of y/x is what separates it from the correctly rounded atan2 (they differ
    in ~20 % of cases); the device code (csrc/k_motion.cu f_atan2) has the same structure."""
    y, x = np.float32(y), np.float32(x)
    z = np.float32(abs(np.float32(y / x)))
    a = np.float32(np.arctan(np.float64(z)))
    return np.float32(-a) if y < 0 else a

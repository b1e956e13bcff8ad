"""Pins oracle/cv_models.py (the numpy specification the CUDA kernels follow) bit-exactly
against the real OpenCV (cv2 4.13.0, setUseOptimized(False)) for every operation the
reference's hot path calls (Stabilizer.cpp:304-305,355-357,449-450,602,611-619,647-649,
740-744,982-987,1056-1060,1121)."""
import numpy as np
import pytest

from oracle import cv_models as M
import synthclip as synth


def _tex(w, h, seed, channels=3):
    img = synth.base_texture(w, h, seed)[synth.MARGIN:-synth.MARGIN, synth.MARGIN:-synth.MARGIN]
    return np.ascontiguousarray(img if channels == 3 else img[..., 1])


def test_bgr2gray(cv2_noopt):
    cv2 = cv2_noopt
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (270, 480, 3), dtype=np.uint8)
    assert np.array_equal(M.bgr2gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


@pytest.mark.parametrize("src,dst,ch", [
    ((1920, 1080), (960, 540), 3),     # exact 2x  -> area fast path      (:449 @1080p)
    ((3840, 2160), (960, 540), 3),     # exact 4x                          (:449 @4K)
    ((1280, 720), (960, 540), 3),      # 1.333x                            (:449 @720p)
    ((1920, 1080), (480, 270), 3),     # first frame                       (:304)
    ((1280, 720), (480, 270), 3),
    ((3840, 2160), (480, 270), 3),
    ((480, 270), (960, 540), 1),       # prevGray up-sampling              (:602)
    ((1860, 1020), (1920, 1080), 3),   # crop+zoom, borderSize=30          (:1121)
    ((1180, 620), (1280, 720), 3),     # crop+zoom, borderSize=50
    ((641, 359), (960, 540), 3),       # odd sizes
])
def test_resize_linear(cv2_noopt, src, dst, ch):
    cv2 = cv2_noopt
    rng = np.random.default_rng(src[0] + dst[0])
    shape = (src[1], src[0], 3) if ch == 3 else (src[1], src[0])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    ref = cv2.resize(img, dst, interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(M.resize_linear(img, dst), ref)


def test_pyr_down(cv2_noopt):
    cv2 = cv2_noopt
    g = _tex(960, 540, 3, 1)
    l1 = M.pyr_down(g)
    assert np.array_equal(l1, cv2.pyrDown(g))
    assert np.array_equal(M.pyr_down(l1), cv2.pyrDown(l1))      # 480x270 -> 240x135 (odd height)
    odd = g[:269, :479]
    assert np.array_equal(M.pyr_down(odd), cv2.pyrDown(odd))


def test_scharr(cv2_noopt):
    cv2 = cv2_noopt
    g = _tex(480, 270, 4, 1)
    ix, iy = M.scharr_deriv(g)
    assert np.array_equal(ix, cv2.Scharr(g, cv2.CV_16S, 1, 0))
    assert np.array_equal(iy, cv2.Scharr(g, cv2.CV_16S, 0, 1))


def _moved_pair(cv2, seed, w=960, h=540, ang=0.3, shift=(3.3, -2.1)):
    big = synth.base_texture(w, h, seed)[..., 1].copy()
    m = cv2.getRotationMatrix2D((big.shape[1] / 2, big.shape[0] / 2), ang, 1.0)
    m[:, 2] += shift
    moved = cv2.warpAffine(big, m, (big.shape[1], big.shape[0]))
    s = synth.MARGIN
    return np.ascontiguousarray(big[s:-s, s:-s]), np.ascontiguousarray(moved[s:-s, s:-s])


@pytest.mark.parametrize("seed,ang,shift", [(5, 0.3, (3.3, -2.1)), (6, -0.5, (-9.5, 6.25))])
def test_lk_bit_exact(cv2_noopt, seed, ang, shift):
    cv2 = cv2_noopt
    prev, nxt = _moved_pair(cv2, seed, ang=ang, shift=shift)
    pts = cv2.goodFeaturesToTrack(prev, 60, 0.02, 15.0, None, blockSize=3).reshape(-1, 2)
    edge = np.array([[0, 0], [959, 539], [3, 200], [958, 10], [500, 538], [1, 1]], np.float32)
    pts = np.vstack([pts, edge])
    ref, st, _ = cv2.calcOpticalFlowPyrLK(
        prev, nxt, pts, None, winSize=(15, 15), maxLevel=2,
        criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 20, 0.03))
    got, gst = M.lk_track(prev, nxt, pts)
    assert np.array_equal(gst, st.ravel())
    ok = st.ravel() == 1
    assert np.array_equal(got[ok].view(np.uint32), ref.reshape(-1, 2)[ok].view(np.uint32))


@pytest.mark.parametrize("w,h", [(960, 540), (480, 270)])
def test_min_eigen_map_bit_exact(cv2_noopt, w, h):
    cv2 = cv2_noopt
    g = _tex(w, h, 7, 1)
    ref = cv2.cornerMinEigenVal(g, 3, ksize=3)
    assert np.array_equal(M.min_eigen_map(g).view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("w,h,mc,q,md", [
    (960, 540, 200, 0.02, 15.0),       # hard-coded re-detection      (:740-744)
    (480, 270, 200, 0.01, 30.0),       # first frame, defaults        (:355-357)
    (960, 540, 50, 0.05, 7.5),
    (480, 270, 0, 0.3, 0.0),           # no cap, no min distance
])
def test_gftt_ordered_list(cv2_noopt, w, h, mc, q, md):
    cv2 = cv2_noopt
    for seed in (1, 2):
        g = _tex(w, h, seed, 1)
        ref = cv2.goodFeaturesToTrack(g, mc, q, md, None, blockSize=3)
        ref = np.zeros((0, 2), np.float32) if ref is None else ref.reshape(-1, 2)
        assert np.array_equal(M.gftt(g, mc, q, md), ref)


def test_gftt_flat_image(cv2_noopt):
    cv2 = cv2_noopt
    g = np.full((270, 480), 77, np.uint8)
    assert cv2.goodFeaturesToTrack(g, 200, 0.01, 30.0, None, blockSize=3) is None
    assert len(M.gftt(g, 200, 0.01, 30.0)) == 0


def test_rng():
    r = M.CvRNG()
    seq = [r.next() for _ in range(4)]
    # multiply-with-carry, seed 2^64-1
    s = 0xFFFFFFFFFFFFFFFF
    exp = []
    for _ in range(4):
        s = ((s & 0xFFFFFFFF) * 4164903690 + (s >> 32)) & 0xFFFFFFFFFFFFFFFF
        exp.append(s & 0xFFFFFFFF)
    assert seq == exp


@pytest.mark.parametrize("n,outl,noise", [(200, 0.0, 0.05), (200, 0.3, 0.2), (60, 0.5, 0.5),
                                          (12, 0.25, 0.1), (4, 0.0, 0.01), (150, 0.7, 0.3)])
def test_ransac_partial_affine(cv2_noopt, n, outl, noise):
    cv2 = cv2_noopt
    for seed in range(8):
        rng = np.random.default_rng(1000 * n + seed)
        src = (rng.random((n, 2)) * (960, 540)).astype(np.float32)
        a = rng.normal(0, 0.01)
        sc = 1 + rng.normal(0, 0.01)
        r = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]) * sc
        dst = src @ r.T + rng.normal(0, 5, 2) + rng.normal(0, noise, (n, 2))
        k = int(outl * n)
        if k:
            dst[:k] += rng.normal(0, 40, (k, 2))
        dst = dst.astype(np.float32)
        ref, mask = cv2.estimateAffinePartial2D(src, dst, None, cv2.RANSAC, 5.0, 500)
        got, gmask = M.estimate_affine_partial_2d(src, dst)
        if ref is None or ref.size == 0:
            assert got is None
            continue
        assert np.array_equal(gmask, mask.ravel())
        assert np.abs(got - ref).max() < 1e-9


@pytest.mark.parametrize("w,h", [(1280, 720), (641, 359)])
def test_warp_affine_bit_exact(cv2_noopt, w, h):
    cv2 = cv2_noopt
    img = _tex(w, h, 9, 3)
    rng = np.random.default_rng(w)
    for i in range(6):
        da = np.float32(rng.normal(0, 0.01))
        t = np.array([[np.cos(da), -np.sin(da), rng.normal(0, 8)],
                      [np.sin(da), np.cos(da), rng.normal(0, 8)]], np.float32)
        if i == 0:
            t = np.array([[1, 0, 0], [0, 1, 0]], np.float32)
        if i == 1:
            t[:, 2] = (w * 1.5, -h * 1.5)                      # fully outside
        ref = cv2.warpAffine(img, t, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
        assert np.array_equal(M.warp_affine(img, t), ref)


@pytest.mark.parametrize("mode,cvname", [(0, "BORDER_CONSTANT"), (1, "BORDER_REPLICATE"), (2, "BORDER_REFLECT"),
                                        (3, "BORDER_WRAP"), (4, "BORDER_REFLECT_101")])
def test_copy_make_border(cv2_noopt, mode, cvname):
    cv2 = cv2_noopt
    img = _tex(320, 180, 11, 3)
    ref = cv2.copyMakeBorder(img, 30, 30, 30, 30, getattr(cv2, cvname), value=(0, 0, 0))
    assert np.array_equal(M.copy_make_border(img, 30, mode), ref)


def test_glibc_cosf_sinf_model_matches_libm():
    """The angle -> matrix step uses glibc's cosf/sinf (not correctly rounded): the model (and the device code that
    follows it) must agree with libm.so.6 bit for bit."""
    import ctypes
    from oracle import cv_models as M
    libm = ctypes.CDLL("libm.so.6")
    for f in (libm.cosf, libm.sinf):
        f.restype, f.argtypes = ctypes.c_float, [ctypes.c_float]
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.normal(0, 0.004, 20000), rng.normal(0, 0.05, 20000), rng.uniform(-0.78, 0.78, 20000),
                         [0.0, 1e-5, -2.0 ** -12, 2.0 ** -12, 0.78]]).astype(np.float32)
    diff_from_correctly_rounded = 0
    for x in xs:
        c, s = M.glibc_cosf_sinf(x)
        assert c == np.float32(libm.cosf(float(x))) and s == np.float32(libm.sinf(float(x))), float(x)
        diff_from_correctly_rounded += int(s != np.float32(np.sin(np.float64(x))))
    assert diff_from_correctly_rounded > 0          # the reason the model exists


def test_add_weighted_model_matches_cv2_simd_path():
    import cv2
    from oracle import cv_models as M
    rng = np.random.default_rng(4)
    a = rng.integers(0, 256, (333, 1001, 3), dtype=np.uint8)
    b = rng.integers(0, 256, (333, 1001, 3), dtype=np.uint8)
    was = cv2.useOptimized()
    cv2.setUseOptimized(True)
    try:
        for alpha in (0.1, 0.25 * (3.0 / 4.0), 0.0, 0.5, 0.0333):
            al = np.float32(alpha)
            be = np.float32(1.0) - al
            ref = cv2.addWeighted(a, float(al), b, float(be), 0.0)
            got = M.add_weighted_u8(a, al, b, be)
            assert int(np.abs(got.astype(int) - ref.astype(int)).max()) <= 1
            # the bulk is the fused SIMD path; only a short scalar tail per thread stripe may round separately
            assert (got != ref).mean() < 1e-4
    finally:
        cv2.setUseOptimized(was)


def test_glibc_atan2f_model_matches_libm():
    import ctypes
    from oracle import cv_models as M
    libm = ctypes.CDLL("libm.so.6")
    libm.atan2f.restype, libm.atan2f.argtypes = ctypes.c_float, [ctypes.c_float, ctypes.c_float]
    rng = np.random.default_rng(5)
    ys = rng.normal(0, 0.004, 40000).astype(np.float32)
    xs = (1 + rng.normal(0, 0.002, 40000)).astype(np.float32)
    same = cr_same = 0
    for y, x in zip(ys, xs):
        r = np.float32(libm.atan2f(float(y), float(x)))
        same += int(M.glibc_atan2f_small(y, x) == r)
        cr_same += int(np.float32(np.arctan2(np.float64(y), np.float64(x))) == r)
    assert same >= len(ys) - 4                   # bit-identical in all but ~1 case per 10^4
    assert cr_same < 0.9 * len(ys)               # the correctly rounded atan2 is NOT what the reference computes


def test_drone_high_freq_chain_known_answers():
    """Hand-traced walk through applyDeadZoneFreeze / applyMicroShakeSuppression / updateTranslationHistory
    (Stabilizer.cpp:2468-2529, 2605-2681) with the default thresholds (2.0 px, 10 frames, 0.9 decay, 1.5 px)."""
    from oracle.stabilizer_ref import Parameters, StabilizerRef
    f32 = np.float32
    st = StabilizerRef(Parameters(droneHighFreqMode=True))
    # 0.5 px: enters the dead zone at once, frozen to zero
    assert np.array_equal(st._hf_filters(np.array([0.5, 0, 0], f32)), np.zeros(3, f32))
    assert st.hf_in_dead_zone and st.hf_freeze_counter == 9
    # 5 px > 1.5 x threshold: leaves the dead zone, accumulator reset, raw motion returned (history too short for a median)
    assert np.array_equal(st._hf_filters(np.array([5, 0, 0], f32)), np.array([5, 0, 0], f32))
    assert not st.hf_in_dead_zone and st.hf_accumulator == 0
    # 2.5 px: outside the dead zone, between shake and 2 x shake of the (zero) median -> 5 % residual
    out = st._hf_filters(np.array([2.5, 0, 0], f32))
    assert out[0] == f32(2.5) * f32(0.05) and out[1] == 0
    assert len(st.hf_hist) == 3

    st = StabilizerRef(Parameters(droneHighFreqMode=True))
    outs = [st._hf_filters(np.array([0.5, 0.25, 0.001], f32)) for _ in range(11)]
    assert all(np.array_equal(o, np.zeros(3, f32)) for o in outs[:9])
    # 10th call: freeze duration over -> raw motion, pulled to 1 % of its distance from the median (0 after nine zeros)
    assert outs[9][0] == f32(0.5) * f32(0.01) and outs[9][1] == f32(0.25) * f32(0.01) and outs[9][2] == f32(0.001)
    # 11th call: still calm -> frozen again
    assert np.array_equal(outs[10], np.zeros(3, f32)) and st.hf_in_dead_zone
    assert len(st.hf_hist) == 10
    # rotation low-pass only with horizonLock
    st = StabilizerRef(Parameters(droneHighFreqMode=True, horizonLock=True))
    o = st._hf_filters(np.array([9, 9, 0.01], f32))
    assert o[2] == f32(f32(0.2) * f32(0.01))
    # box radius clamp [10, 50] in drone mode
    p = np.arange(40, dtype=f32)
    assert StabilizerRef.box_at(p, 30, 20, True) == f32(19.5) and StabilizerRef.box_at(p, 3, 0, True) == f32(5)
    assert StabilizerRef.box_at(p, 3, 0, False) == f32(1.5)

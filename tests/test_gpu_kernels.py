"""Per-kernel parity through the C-ABI (vs_k_*) against the oracle (cv2 4.13 on its baseline
path + oracle/cv_models.py), on seeded inputs.  Integer/byte/index work is compared bit-exactly."""
import numpy as np

import synthclip
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def vsb():
    import __graft_entry__
    __graft_entry__.build()
    import video_stab_b200
    assert torch.cuda.is_available()
    return video_stab_b200


def _tex(vsb, w, h, seed, ch=3):
    m = synthclip.MARGIN
    img = synthclip.base_texture(w, h, seed)[m:-m, m:-m]
    return np.ascontiguousarray(img if ch == 3 else img[..., 1])


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _matrices(n, seed, scale=8.0):
    rng = np.random.default_rng(seed)
    T = np.zeros((n, 2, 3), np.float32)
    for i in range(n):
        da = np.float32(rng.normal(0, 0.01))
        T[i] = [[np.cos(da), -np.sin(da), rng.normal(0, scale)], [np.sin(da), np.cos(da), rng.normal(0, scale)]]
    T[0] = [[1, 0, 0], [0, 1, 0]]
    return T


@pytest.mark.parametrize("w,h", [(1280, 720), (1920, 1080), (641, 359)])
def test_warp_affine_bit_exact(vsb, cv2_noopt, w, h):
    cv2 = cv2_noopt
    n = 4
    frames = np.stack([_tex(vsb, w, h, 20 + i) for i in range(n)])
    T = _matrices(n, w)
    T[1, :, 2] = (w * 1.5, -h * 1.5)          # completely outside -> all black
    T[2, :, 2] = (-37.25, 21.5)               # large shift: wide zero band
    out = vsb.kernels.warp_affine(_dev(frames), T).cpu().numpy()
    for i in range(n):
        ref = cv2.warpAffine(frames[i], T[i], (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
        assert np.array_equal(out[i], ref), f"frame {i}: max diff {np.abs(out[i].astype(int) - ref).max()}"


@pytest.mark.parametrize("w,h", [(1920, 1080), (644, 362)])
def test_warp_affine_general_matrices(vsb, cv2_noopt, w, h):
    """Rotations from 0.5 to 60 degrees, scales, flips, shears and far translations: exercises the widest staged
    boxes, the per-tile generic fallback and tiles that lie partly or wholly outside the source."""
    cv2 = cv2_noopt
    f = _tex(vsb, w, h, 77)
    rng = np.random.default_rng(w)
    mats = []
    for ang in (0.009, 0.02, 0.03, 0.05, 0.069, 0.075, 0.2, 1.05, -0.4, 3.1):
        for sc in (1.0, 0.8, 1.3):
            c, s = np.cos(ang) * sc, np.sin(ang) * sc
            mats.append([[c, -s, rng.normal(0, 40)], [s, c, rng.normal(0, 40)]])
    mats += [[[-1, 0, w - 1], [0, 1, 0]], [[1, 0.2, -30], [0.1, 1, 5]], [[1, 0, 0.5], [0, 1, 0.5]],
             [[1, 0, -w + 3], [0, 1, 2.25]], [[1, 0, 5.03125], [0, 1, h - 2]], [[0.25, 0, 0], [0, 0.25, 0]],
             [[4, 0, -w], [0, 4, -h]], [[1e-3, 0, 10], [0, 1e-3, 10]], [[1, 0, 1e6], [0, 1, -1e6]]]
    T = np.asarray(mats, np.float32)
    n = len(T)
    out = vsb.kernels.warp_affine(_dev(np.broadcast_to(f, (n,) + f.shape)), T).cpu().numpy()
    for i in range(n):
        ref = cv2.warpAffine(f, T[i], (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
        assert np.array_equal(out[i], ref), f"matrix {i} {T[i].tolist()}: max diff {np.abs(out[i].astype(int) - ref).max()}"


@pytest.mark.parametrize("w,h,n", [(1920, 1080, 24), (1284, 722, 12), (256, 301, 12), (3840, 2160, 3)])
def test_warp_quad_paths_small_angles(vsb, cv2_noopt, w, h, n):
    """k_warp_quad (four adjacent pixels per lane, taps straight from the raw TMA box): a sweep over the angles a stabiliser
    applies (1e-4 .. 0.06 rad, both signs, scales a hair off 1) with sub-pixel shifts, on sizes whose last tile is short and
    on batches long enough for 8-tile strips.  Exercises all four byte phases, quads with an adelta step inside, quads that
    straddle a change of source row, steps where Y0 does not advance by 1024 per row, and the per-pixel step path."""
    cv2 = cv2_noopt
    rng = np.random.default_rng(w * 7 + h)
    f = _tex(vsb, w, h, 91)
    angs = np.concatenate([np.geomspace(1e-4, 0.06, n - 2), [0.0, 0.0035]]) * rng.choice([-1.0, 1.0], n)
    T = np.zeros((n, 2, 3), np.float32)
    for i, a in enumerate(angs):
        sc = 1.0 + (rng.normal(0, 2e-4) if i % 3 == 0 else 0.0)
        T[i] = [[np.cos(a) * sc, -np.sin(a) * sc, rng.normal(0, 6)], [np.sin(a) * sc, np.cos(a) * sc, rng.normal(0, 6)]]
    out = vsb.kernels.warp_affine(_dev(np.broadcast_to(f, (n,) + f.shape)), T).cpu().numpy()
    for i in range(n):
        ref = cv2.warpAffine(f, T[i], (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
        assert np.array_equal(out[i], ref), f"matrix {i} {T[i].tolist()}: max diff {np.abs(out[i].astype(int) - ref).max()}"


def test_warp_affine_unaligned_views(vsb, cv2_noopt):
    """Source / destination views whose base address or row stride is not 4-byte aligned (ROI of a larger frame)."""
    cv2 = cv2_noopt
    big = _tex(vsb, 700, 400, 78)
    T = _matrices(3, 3)[1:2]
    for x0, ww in ((1, 641), (3, 640), (2, 322)):
        view = big[5:365, x0:x0 + ww]
        d = _dev(big)[5:365, x0:x0 + ww]
        out = vsb.kernels.warp_affine(d[None], T).cpu().numpy()[0]
        ref = cv2.warpAffine(np.ascontiguousarray(view), T[0], (ww, 360), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
        assert np.array_equal(out, ref)


def test_warp_affine_4k_and_linearity_property(vsb, cv2_noopt):
    cv2 = cv2_noopt
    w, h = 3840, 2160
    f = _tex(vsb, w, h, 31)
    T = _matrices(2, 5)
    out = vsb.kernels.warp_affine(_dev(np.stack([f, f])), T).cpu().numpy()
    assert np.array_equal(out[0], f)                               # identity is the identity
    ref = cv2.warpAffine(f, T[1], (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
    assert np.array_equal(out[1], ref)
    # size-independent property: integer translation == shifted copy with a zero band
    Tt = np.array([[[1, 0, 7], [0, 1, -3]]], np.float32)
    o = vsb.kernels.warp_affine(_dev(f[None]), Tt).cpu().numpy()[0]
    exp = np.zeros_like(f)
    exp[:-3, 7:] = f[3:, :-7]
    assert np.array_equal(o, exp)


@pytest.mark.parametrize("mode,bmode,cvname", [(1, 0, "BORDER_CONSTANT"), (1, 1, "BORDER_REPLICATE"),
                                              (1, 2, "BORDER_REFLECT"), (1, 3, "BORDER_WRAP"),
                                              (1, 4, "BORDER_REFLECT_101")])
def test_border_then_warp(vsb, cv2_noopt, mode, bmode, cvname):
    """copyMakeBorder + warpAffine (Stabilizer.cpp:981-990, 1056-1060).  (640, 360) and (1920, 1080) take the tiled
    TMA kernel (interior tiles staged, margin tiles per pixel unless the margin is constant 0), (642, 360) the
    per-pixel kernel (row pitch not a multiple of 16 bytes)."""
    cv2 = cv2_noopt
    for w, h, b in ((640, 360, 24), (642, 360, 24), (1920, 1080, 100)):
        f = _tex(vsb, w, h, 41)
        Ts = [_matrices(3, 9)[2], np.array([[1, 0, 0], [0, 1, 0]], np.float32)]
        a = np.deg2rad(12.0)
        Ts.append(np.array([[np.cos(a), -np.sin(a), 40.5], [np.sin(a), np.cos(a), -60.25]], np.float32))
        for k, T in enumerate(Ts[:2] if w == 1920 else Ts):
            out = vsb.kernels.warp_output(_dev(f), T, mode, b, bmode).cpu().numpy()
            src = cv2.copyMakeBorder(f, b, b, b, b, getattr(cv2, cvname), value=(0, 0, 0))
            ref = cv2.warpAffine(src, T, (w + 2 * b, h + 2 * b), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
            assert np.array_equal(out, ref), f"{w}x{h} matrix {k}: {int((out != ref).sum())} bytes differ"


@pytest.mark.parametrize("w,h,b", [(1280, 720, 30), (1920, 1080, 50), (3840, 2160, 30), (1920, 1080, 1), (1920, 1080, 300),
                                   (1000, 562, 17), (640, 360, 8), (136, 64, 2)])
def test_warp_crop_zoom(vsb, cv2_noopt, w, h, b):
    """cropNZoom output stage (Stabilizer.cpp:1056-1060, 1108-1124) = warpAffine, crop by b, cv::resize back.  Frames
    whose rows are 16-byte aligned take the fused single-pass kernel, (1000, 562) the two-pass fallback; the matrices
    cover the identity, typical jitter, a shift that leaves a zero band, and a 25 degree rotation (per-tile generic
    path inside the fused kernel)."""
    cv2 = cv2_noopt
    f = _tex(vsb, w, h, 43)
    Ts = list(_matrices(3, 11))
    Ts.append(np.array([[1, 0, -37.25], [0, 1, 21.5]], np.float32))
    a = np.deg2rad(25.0)
    Ts.append(np.array([[np.cos(a), -np.sin(a), 0.1 * w], [np.sin(a), np.cos(a), -0.2 * h]], np.float32))
    Ts.append(np.array([[1.07, 0.02, 3.5], [-0.03, 0.95, -2.25]], np.float32))
    for k, T in enumerate(Ts):
        out = vsb.kernels.warp_output(_dev(f), T, 2, b, 0).cpu().numpy()
        wr = cv2.warpAffine(f, T, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
        ref = cv2.resize(wr[b:h - b, b:w - b].copy(), (w, h))
        assert np.array_equal(out, ref), f"matrix {k}: {int((out != ref).sum())} bytes differ, max {np.abs(out.astype(int) - ref).max()}"


@pytest.mark.parametrize("w,h", [(1920, 1080), (1280, 720), (3840, 2160), (1000, 562)])
def test_gray_pyramid_bit_exact(vsb, cv2_noopt, w, h):
    cv2 = cv2_noopt
    f = _tex(vsb, w, h, 51)
    lv = [t.cpu().numpy() for t in vsb.kernels.gray_pyramid(_dev(f))]
    g = cv2.cvtColor(cv2.resize(f, (960, 540), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY)
    assert np.array_equal(lv[0], g)
    g1 = cv2.pyrDown(g)
    assert np.array_equal(lv[1], g1)
    assert np.array_equal(lv[2], cv2.pyrDown(g1))
    small = vsb.kernels.gray_pyramid(_dev(f), first_frame=True)[0].cpu().numpy()
    assert np.array_equal(small, cv2.cvtColor(cv2.resize(f, (480, 270), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2GRAY))


@pytest.mark.parametrize("src,dst,ch", [((480, 270), (960, 540), 1), ((1860, 1020), (1920, 1080), 3),
                                        ((1920, 1080), (960, 540), 3), ((641, 359), (960, 540), 3)])
def test_resize_linear_bit_exact(vsb, cv2_noopt, src, dst, ch):
    cv2 = cv2_noopt
    rng = np.random.default_rng(src[0])
    shape = (src[1], src[0], 3) if ch == 3 else (src[1], src[0])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    out = vsb.kernels.resize_linear(_dev(img), dst).cpu().numpy()
    assert np.array_equal(out, cv2.resize(img, dst, interpolation=cv2.INTER_LINEAR))


@pytest.mark.parametrize("w,h,mc,q,md", [(960, 540, 200, 0.02, 15.0), (480, 270, 200, 0.01, 30.0),
                                         (960, 540, 50, 0.05, 7.5), (480, 270, 0, 0.3, 0.0),
                                         (960, 540, 1000, 0.001, 3.0), (480, 270, 300, 0.01, 10.0)])
def test_good_features_ordered_list_bit_exact(vsb, cv2_noopt, w, h, mc, q, md):
    cv2 = cv2_noopt
    for seed in (1, 2, 3):
        g = _tex(vsb, w, h, seed, 1)
        ref = cv2.goodFeaturesToTrack(g, mc, q, md, None, blockSize=3)
        ref = np.zeros((0, 2), np.float32) if ref is None else ref.reshape(-1, 2)
        got = vsb.kernels.good_features(_dev(g), mc, q, md)
        assert got.shape == ref.shape, (got.shape, ref.shape)
        assert np.array_equal(got, ref)


def test_good_features_flat_and_sparse(vsb, cv2_noopt):
    cv2 = cv2_noopt
    flat = np.full((540, 960), 90, np.uint8)
    assert len(vsb.kernels.good_features(_dev(flat), 200, 0.02, 15.0)) == 0
    one = flat.copy()
    one[200:230, 300:340] = 200                       # a single bright rectangle: 4 corners
    ref = cv2.goodFeaturesToTrack(one, 200, 0.02, 15.0, None, blockSize=3).reshape(-1, 2)
    assert np.array_equal(vsb.kernels.good_features(_dev(one), 200, 0.02, 15.0), ref)


def test_good_features_massive_ties(vsb, cv2_noopt):
    """Checkerboards give thousands of candidates with IDENTICAL eigenvalues: exercises the tie order
    (higher address first), the multi-chunk path and the single-bin overflow split of k_select."""
    cv2 = cv2_noopt
    yy, xx = np.mgrid[0:540, 0:960]
    for sq, md, mc in ((8, 3.0, 0), (8, 0.0, 0), (12, 15.0, 200), (6, 1.0, 1500)):
        img = (((yy // sq) + (xx // sq)) % 2 * 160 + 40).astype(np.uint8)
        ref = cv2.goodFeaturesToTrack(img, mc, 0.02, md, None, blockSize=3)
        ref = np.zeros((0, 2), np.float32) if ref is None else ref.reshape(-1, 2)
        got = vsb.kernels.good_features(_dev(img), mc, 0.02, md)
        n = min(len(got), 2048)                 # the kernel caps "unlimited" at 2048 corners
        assert len(ref) >= n and (len(got) == len(ref) or len(got) == 2048)
        assert np.array_equal(got[:n], ref[:n]), f"square {sq} minDist {md}"


def _moved_pair(vsb, cv2, seed, ang, shift, w=960, h=540):
    big = synthclip.base_texture(w, h, seed)[..., 1].copy()
    m = cv2.getRotationMatrix2D((big.shape[1] / 2, big.shape[0] / 2), ang, 1.0)
    m[:, 2] += shift
    moved = cv2.warpAffine(big, m, (big.shape[1], big.shape[0]))
    s = synthclip.MARGIN
    return np.ascontiguousarray(big[s:-s, s:-s]), np.ascontiguousarray(moved[s:-s, s:-s])


@pytest.mark.parametrize("seed,ang,shift", [(5, 0.3, (3.3, -2.1)), (6, -0.5, (-9.5, 6.25)), (7, 0.0, (25.0, 14.0))])
def test_pyr_lk_bit_exact(vsb, cv2_noopt, seed, ang, shift):
    cv2 = cv2_noopt
    prev, nxt = _moved_pair(vsb, cv2, seed, ang, shift)
    pts = cv2.goodFeaturesToTrack(prev, 200, 0.02, 15.0, None, blockSize=3).reshape(-1, 2)
    edge = np.array([[0, 0], [959, 539], [3, 200], [958, 10], [500, 538], [1, 1], [480.5, 270.25]], np.float32)
    pts = np.vstack([pts, edge])
    ref, st, _ = cv2.calcOpticalFlowPyrLK(prev, nxt, pts, None, winSize=(15, 15), maxLevel=2,
                                          criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 20, 0.03))
    got, gst = vsb.kernels.pyr_lk(_dev(prev), _dev(nxt), pts)
    assert np.array_equal(gst, st.ravel())
    ok = st.ravel() == 1
    assert np.array_equal(got[ok].view(np.uint32), ref.reshape(-1, 2)[ok].view(np.uint32)), \
        f"max diff {np.abs(got[ok] - ref.reshape(-1, 2)[ok]).max()}"


@pytest.mark.parametrize("seed,ang,shift", [(81, 0.1, (14.5, 12.25)), (82, 0.0, (20.0, 18.0)), (83, -0.2, (-16.0, -13.5))])
def test_pyr_lk_points_leaving_the_frame(vsb, cv2_noopt, seed, ang, shift):
    """Points along the four borders with a motion that carries many of them out of the frame: OpenCV drops the status
    of a point whose FINAL window origin lies outside level 0 (the re-check it makes when it computes `err`, which the
    reference requests), besides the checks inside the iteration loop."""
    cv2 = cv2_noopt
    prev, nxt = _moved_pair(vsb, cv2, seed, ang, shift)
    xs = np.arange(8, 960, 37, dtype=np.float32)
    ys = np.arange(10, 540, 38, dtype=np.float32)
    pts = np.concatenate([np.stack([xs, np.full_like(xs, 536.0)], 1), np.stack([xs, np.full_like(xs, 2.0)], 1),
                          np.stack([np.full_like(ys, 956.0), ys], 1), np.stack([np.full_like(ys, 3.0), ys], 1)]).astype(np.float32)
    ref, st, _ = cv2.calcOpticalFlowPyrLK(prev, nxt, pts, None, winSize=(15, 15), maxLevel=2,
                                          criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 20, 0.03))
    ref, st = ref.reshape(-1, 2), st.ravel()
    org = np.floor(ref - 7)
    left = (org[:, 0] < -15) | (org[:, 0] >= 960) | (org[:, 1] < -15) | (org[:, 1] >= 540)
    assert int((left & (st == 0)).sum()) >= 5, "the case no longer exercises points that end outside the frame"
    got, gst = vsb.kernels.pyr_lk(_dev(prev), _dev(nxt), pts)
    assert np.array_equal(gst, st), f"status differs at {np.nonzero(gst != st)[0]}"
    ok = st == 1
    assert np.array_equal(got[ok].view(np.uint32), ref[ok].view(np.uint32))


@pytest.mark.parametrize("n,outl,noise", [(200, 0.0, 0.05), (200, 0.3, 0.2), (60, 0.5, 0.5), (12, 0.25, 0.1),
                                          (4, 0.0, 0.01), (150, 0.7, 0.3), (1500, 0.4, 0.3), (3, 0.0, 0.0), (0, 0, 0)])
def test_ransac_partial_affine(vsb, cv2_noopt, n, outl, noise):
    cv2 = cv2_noopt
    for seed in range(6):
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
        got, gmask, iters = vsb.kernels.estimate_affine_partial(src, dst)
        if n < 4:                       # the reference only calls the estimator with >= 4 pairs (:645)
            assert got is None
            continue
        ref, mask = cv2.estimateAffinePartial2D(src, dst, None, cv2.RANSAC, 5.0, 500)
        if ref is None or ref.size == 0:
            assert got is None
            continue
        assert got is not None
        assert np.array_equal(gmask, mask.ravel()), "inlier mask differs"
        # 1e-3 px of corner displacement (north star); the fit itself agrees to ~1e-9
        assert np.abs(got - ref).max() < 1e-8


@pytest.mark.parametrize("w,h,block,q,md", [(480, 270, 5, 0.01, 30.0), (480, 270, 7, 0.02, 10.0), (480, 270, 2, 0.01, 15.0),
                                            (960, 540, 5, 0.05, 12.0), (480, 270, 9, 0.01, 8.0), (480, 270, 1, 0.03, 20.0)])
def test_good_features_other_block_sizes(vsb, cv2_noopt, w, h, block, q, md):
    """cv::goodFeaturesToTrack with blockSize != 3 (the first-frame detection passes params.blockSize, Stabilizer.cpp:355-357):
    ordered corner list bit-exact against cv2, odd and even window sizes."""
    cv2 = cv2_noopt
    g = _tex(vsb, w, h, 40 + block, 1)
    ref = cv2.goodFeaturesToTrack(g, 200, q, md, None, blockSize=block)
    ref = np.zeros((0, 2), np.float32) if ref is None else ref.reshape(-1, 2)
    got = vsb.kernels.good_features(_dev(g), 200, q, md, block_size=block)
    assert len(ref) > 20
    assert np.array_equal(got, ref), f"first difference at {next((i for i in range(min(len(got), len(ref))) if not np.array_equal(got[i], ref[i])), None)} of {len(ref)} / {len(got)}"


@pytest.mark.parametrize("w,h", [(1920, 1080), (1280, 720), (3840, 2160), (642, 362), (64, 2)])
def test_nv12_bgr_conversions_bit_exact(vsb, cv2_noopt, w, h):
    """Decoder / encoder hand-off (vs_nv12_to_bgr_device, vs_bgr_to_nv12_device): bit-exact with cv2's NV12 -> BGR and
    BGR -> I420 (U, V interleaved), on random bytes (every saturation case), aligned and unaligned widths / views."""
    cv2 = cv2_noopt
    rng = np.random.default_rng(w + h)
    nv = rng.integers(0, 256, (h * 3 // 2, w), dtype=np.uint8)
    nv[:8, :8] = 0
    nv[8:16, :8] = 255
    ref = cv2.cvtColor(nv, cv2.COLOR_YUV2BGR_NV12)
    got = vsb.kernels.nv12_to_bgr(_dev(nv)).cpu().numpy()
    assert np.array_equal(got, ref), np.abs(got.astype(int) - ref).max()
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    i420 = cv2.cvtColor(img, cv2.COLOR_BGR2YUV_I420)
    flat, q = i420.reshape(-1), (h // 2) * (w // 2)
    u = flat[h * w:h * w + q].reshape(h // 2, w // 2)
    v = flat[h * w + q:].reshape(h // 2, w // 2)
    ref_nv = np.concatenate([i420[:h], np.stack([u, v], -1).reshape(h // 2, w)], 0)
    got_nv = vsb.kernels.bgr_to_nv12(_dev(img)).cpu().numpy()
    assert np.array_equal(got_nv, ref_nv), np.abs(got_nv.astype(int) - ref_nv).max()
    if w > 200:
        # views into larger surfaces: pitch 2048-style strides and an odd byte offset (the per-block path)
        big = torch.zeros((h * 3 // 2, w + 37), dtype=torch.uint8, device="cuda")
        big[:, 5:5 + w] = _dev(nv)
        view = big[:, 5:5 + w]
        out = torch.zeros((h, w + 11, 3), dtype=torch.uint8, device="cuda")
        vsb.kernels.nv12_to_bgr(view, out=out[:, 3:3 + w])
        assert np.array_equal(out[:, 3:3 + w].cpu().numpy(), ref)
        back = torch.zeros((h * 3 // 2, w + 37), dtype=torch.uint8, device="cuda")
        vsb.kernels.bgr_to_nv12(_dev(img), out=back[:, 4:4 + w])
        assert np.array_equal(back[:, 4:4 + w].cpu().numpy(), ref_nv)


def test_nv12_round_trip_through_the_stabilizer(vsb):
    """NV12 surface -> BGR on the device -> vs_stabilizer_push_device (borrowed) -> NV12, stream-ordered with events only:
    the decode -> stabilize -> encode shape of INTEGRATION.md.  The BGR frames the stabilizer sees equal the host-converted ones,
    so its outputs equal those of the host-fed handle."""
    import cv2
    w, h, n = 640, 360, 20
    clip = synthclip.make_clip(w, h, n, 99)
    nvs = []
    for f in clip:
        i420 = cv2.cvtColor(f, cv2.COLOR_BGR2YUV_I420)
        u = i420[h:h + h // 4].reshape(h // 2, w // 2)
        v = i420[h + h // 4:].reshape(h // 2, w // 2)
        nvs.append(np.concatenate([i420[:h], np.stack([u, v], -1).reshape(h // 2, w)], 0))
    params = vsb.Parameters(smoothingRadius=5)
    ref_st = vsb.Stabilizer(params)
    ref_out = []
    for nv in nvs:
        o = ref_st.stabilize(cv2.cvtColor(nv, cv2.COLOR_YUV2BGR_NV12))
        ref_out.append(None if o is None else o.copy())
    st = vsb.Stabilizer(params)
    dec = torch.cuda.Stream()
    d_nv = [_dev(nv) for nv in nvs]
    bgr = torch.empty((n, h, w, 3), dtype=torch.uint8, device="cuda")
    outs = torch.empty((n, h, w, 3), dtype=torch.uint8, device="cuda")
    enc = torch.empty((n, h * 3 // 2, w), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ext = torch.cuda.ExternalStream(st.stream)
    produced = []
    for k in range(n):
        vsb.kernels.nv12_to_bgr(d_nv[k], out=bgr[k], stream=dec.cuda_stream)
        ev = torch.cuda.Event()
        ev.record(dec)
        st.wait_event(ev.cuda_event)
        got = st.push_device(bgr[k].data_ptr(), w, h, w * 3, outs[k].data_ptr(), w * 3, h * w * 3, borrow=True)
        if got is not None:
            vsb.kernels.bgr_to_nv12(outs[k], out=enc[k], stream=st.stream)
            produced.append(k)
    st.sync()
    assert produced == [k for k in range(n) if ref_out[k] is not None]
    for k in produced:
        assert np.array_equal(outs[k].cpu().numpy(), ref_out[k])
        i420 = cv2.cvtColor(ref_out[k], cv2.COLOR_BGR2YUV_I420)
        assert np.array_equal(enc[k, :h].cpu().numpy(), i420[:h])

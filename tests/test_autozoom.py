"""vs::AutoZoomCrop.  The host half (border following + the rectangle-shrinking loop, which the
reference also runs on the CPU) is checked here without a GPU: contours bit-equal to cv2.findContours on random masks, crop result
bit-equal to the REFERENCE'S OWN AutoZoomCrop.cpp (oracle/_ref/libstages_ref.so).  The `-m gpu` tests run the whole call on the
device (gray, threshold, 5x5 elliptic close, crop + scale) against the same compiled reference."""
import numpy as np
import pytest

import synthclip


@pytest.fixture(scope="module")
def vsb():
    import __graft_entry__
    __graft_entry__.build()
    import video_stab_b200
    return video_stab_b200


def _ref_zoom():
    from oracle import ref_stages
    if not ref_stages.available():
        pytest.skip("oracle/_ref/libstages_ref.so not present")
    return ref_stages.RefAutoZoomCrop()


def _random_mask(rng, kind, w, h):
    import cv2
    if kind == 0:
        return (rng.random((h, w)) < rng.uniform(0.2, 0.8)).astype(np.uint8) * 255
    if kind == 1:
        m = np.zeros((h, w), np.uint8)
        for _ in range(int(rng.integers(1, 6))):
            cv2.ellipse(m, (int(rng.integers(0, w)), int(rng.integers(0, h))), (int(rng.integers(1, w // 2 + 2)), int(rng.integers(1, h // 2 + 2))),
                        float(rng.uniform(0, 180)), 0, 360, 255, -1)
        if rng.random() < 0.5:
            m[rng.random((h, w)) < 0.05] = 0
        return m
    if kind == 2:
        m = np.full((h, w), 255, np.uint8)
        m[rng.random((h, w)) < 0.1] = 0
        return m
    return cv2.morphologyEx((rng.random((h, w)) < 0.5).astype(np.uint8) * 255, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (5, 5)))


def test_external_contours_equal_cv2_find_contours(vsb):
    """RETR_EXTERNAL + CHAIN_APPROX_SIMPLE: same contours, same vertices, same order as cv2 on 240 random masks (noise, blobs with
    holes and specks, full frames with pinholes, closed noise) incl. 1-pixel components and regions touching the image border."""
    import cv2
    rng = np.random.default_rng(0)
    for trial in range(240):
        w, h = int(rng.integers(1, 90)), int(rng.integers(1, 70))
        m = np.ascontiguousarray(_random_mask(rng, trial % 4, w, h))
        ref, _ = cv2.findContours(m.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        got = vsb.AutoZoomCrop.find_external_contours(m)
        assert len(ref) == len(got), f"trial {trial}: {len(got)} contours vs {len(ref)}"
        for i, (a, b) in enumerate(zip(ref, got)):
            assert np.array_equal(a.reshape(-1, 2), b), f"trial {trial}: contour {i}"
    assert vsb.AutoZoomCrop.find_external_contours(np.zeros((5, 7), np.uint8)) == []
    assert vsb.AutoZoomCrop.rect_from_mask(np.zeros((5, 7), np.uint8)) is None


ZOOM_CASES = [(1280, 720), (1920, 1080), (640, 360), (1001, 563)]


def _zoom_frames(n):
    rng = np.random.default_rng(1)
    for trial in range(n):
        w, h = ZOOM_CASES[trial % 4]
        ang = float(rng.uniform(-8, 8))
        shift = rng.uniform(-30, 30, 2)
        yield trial, synthclip.black_corner_frame(w, h, trial, ang, shift, hole=trial % 6 == 5)


def test_crop_rectangle_equals_reference_cpp(vsb, cv2_noopt):
    """24 random (angle, shift) pairs, four frame sizes, some with a black object inside the content: the host half fed with
    cv2's closed content mask gives exactly the frame the compiled reference returns (crop rectangle bit-equal)."""
    cv2 = cv2_noopt
    za = _ref_zoom()
    for trial, fr in _zoom_frames(24):
        want = za.crop(fr)
        gray = cv2.cvtColor(fr, cv2.COLOR_BGR2GRAY)
        cm = cv2.threshold(gray, 1, 255, cv2.THRESH_BINARY)[1]
        cm = cv2.morphologyEx(cm, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (5, 5)))
        r = vsb.AutoZoomCrop.rect_from_mask(cm)
        assert r is not None and r[2] > 0 and r[3] > 0
        x, y, rw, rh = r
        T = np.array([[640.0 / rw, 0, 0], [0, 360.0 / rh, 0]], np.float32)
        got = cv2.warpAffine(np.ascontiguousarray(fr[y:y + rh, x:x + rw]), T, (640, 360), flags=cv2.INTER_LINEAR)
        assert want.shape == got.shape and np.array_equal(want, got), f"trial {trial}: rect {r}"


@pytest.mark.gpu
def test_content_mask_kernel_equals_cv2(vsb, cv2_noopt):
    import ctypes as C

    import torch
    from video_stab_b200._capi import lib
    cv2 = cv2_noopt
    for trial, fr in _zoom_frames(8):
        h, w = fr.shape[:2]
        fr = fr.copy()
        fr[5:9, 5:9] = 0                                   # specks the close must fill
        fr[h // 2, ::7] = 0
        d = torch.from_numpy(fr).cuda()
        m = torch.empty((h, w), dtype=torch.uint8, device="cuda")
        s = torch.empty_like(m)
        assert lib.vs_k_content_mask(d.data_ptr(), w, h, w * 3, m.data_ptr(), s.data_ptr(), None) == 0
        torch.cuda.synchronize()
        cm = cv2.threshold(cv2.cvtColor(fr, cv2.COLOR_BGR2GRAY), 1, 255, cv2.THRESH_BINARY)[1]
        cm = cv2.morphologyEx(cm, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (5, 5)))
        assert np.array_equal(m.cpu().numpy(), cm), f"trial {trial}"


@pytest.mark.gpu
def test_auto_zoom_crop_vs_reference(vsb):
    """the whole call (host API and device API) against the compiled reference: 24 frames, every output pixel"""
    import torch
    za = _ref_zoom()
    for trial, fr in _zoom_frames(24):
        want = za.crop(fr)
        got = vsb.AutoZoomCrop.autoZoomCrop(fr)
        assert got.shape == want.shape == (360, 640, 3)
        assert np.array_equal(got, want), f"trial {trial}: {np.abs(got.astype(int) - want.astype(int)).max()} LSB"
        h, w = fr.shape[:2]
        d = torch.from_numpy(fr).cuda()
        o = torch.zeros((360, 640, 3), dtype=torch.uint8, device="cuda")
        assert vsb.AutoZoomCrop.crop_device(d.data_ptr(), w, h, w * 3, o.data_ptr(), 640 * 3, 640 * 360 * 3) == (640, 360)
        torch.cuda.synchronize()
        assert np.array_equal(o.cpu().numpy(), want)
    # a frame with no content at all: no contour -> the frame comes back unchanged
    black = np.zeros((360, 640, 3), np.uint8)
    assert np.array_equal(vsb.AutoZoomCrop.autoZoomCrop(black), za.crop(black))


@pytest.mark.gpu
def test_roll_then_stabilize_then_zoom_pipeline(vsb):
    """The reference's frame pipeline order (examples/vsg.cpp:1272-1285, roll-correction-file.cpp:61-66): roll correction ->
    stabilize -> auto zoom-crop, every stage against its compiled reference, chained on the reference's own outputs."""
    from oracle import ref_lib, ref_stages
    from oracle.stabilizer_ref import Parameters
    if not (ref_stages.available() and ref_lib.available()):
        pytest.skip("oracle/_ref not present")
    w, h, n = 1280, 720, 14
    clip = synthclip.horizon_clip(w, h, n, 33)
    kw = dict(angleFilterMin=-70.0, angleFilterMax=70.0)
    r_roll, r_stab, r_zoom = ref_stages.RefRollCorrection(ref_stages.RollParameters(**kw)), ref_lib.RefStabilizer(Parameters(smoothingRadius=5)), ref_stages.RefAutoZoomCrop()
    roll, stab = vsb.RollCorrection(vsb.RollParameters(**kw)), vsb.Stabilizer(vsb.Parameters(smoothingRadius=5))
    outs = 0
    for f in clip:
        a, b = r_roll.correct(f), roll.autoCorrectRoll(f)
        assert np.array_equal(a, b)
        sa, sb = r_stab.stabilize(a), stab.stabilize(b)
        assert (sa is None) == (sb is None)
        if sa is not None:
            d = np.abs(sa.astype(np.int16) - sb.astype(np.int16))
            assert d[40:-40, 40:-40].max() <= 1
            za, zb = r_zoom.crop(sa), vsb.AutoZoomCrop.autoZoomCrop(sa)
            assert np.array_equal(za, zb)
            outs += 1
    assert outs >= n - 5

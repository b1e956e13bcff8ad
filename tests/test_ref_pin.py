"""The parity pin: the oracle against the REFERENCE ITSELF.

`oracle/_ref/libvideostab_ref.so` is /root/reference/src/Stabilizer.cpp — all of it, unmodified — compiled by
`oracle/build_ref.py` against a header stand-in whose image operations call the real OpenCV (cv2 4.13).  These
tests run the reference's own `vs::Stabilizer` on seeded clips and require the Python restatement
(`oracle/stabilizer_ref.py`, what the GPU tests and the goldens are built on) to agree with it BIT FOR BIT: every
per-frame transform and path sample, every corner list, LK status vector and RANSAC mask, the smoothed path sample,
adaptive radius and motion intent of every emitted frame, the float32 warp matrix, and every output pixel.  The pure
host functions (`boxFilterConvolve` Stabilizer.cpp:1139-1172, `gaussianFilterConvolve` :1364-1413, `kalmanFilterSmooth`
:1416-1458, `adaptSmoothingRadius` :1461-1492, `calculateAdaptiveRadius` :1637-1673, `analyzeMotionIntent` :1676-1719,
`calculateAdaptiveStabilizationStrength` :1722-1747, variance / consistency :1750-1780, the drone chain :2447-2686) are
also driven on their own with random inputs.  The committed goldens are re-derived from the reference as well.

CPU only.  Where neither /root/reference nor a prebuilt oracle/_ref exists the module is skipped."""
import os
import zlib

import numpy as np
import pytest

import synthclip
from oracle import ref_lib
from oracle.stabilizer_ref import Parameters, StabilizerRef, run_clip

pytestmark = pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref not built and /root/reference absent")

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
u32 = np.uint32


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(u32)


def assert_same_run(clip, params):
    """reference C++ vs Python restatement on one clip: everything, bit for bit."""
    o_port, port = run_clip(clip, params)
    o_ref, ref = ref_lib.run_clip(clip, params)
    assert not ref.ops.errors, ref.ops.errors[:3]
    assert len(o_port) == len(o_ref)
    assert np.array_equal(port.first_corners, ref.first_corners)
    assert np.array_equal(bits(np.array(port.transforms).reshape(-1, 3)), bits(ref.transforms()))
    assert np.array_equal(bits(np.array(port.path).reshape(-1, 3)), bits(ref.path()))
    assert len(port.frame_records) == len(ref.frame_records)
    for a, b in zip(port.frame_records, ref.frame_records):
        assert a.frame_index == b.frame_index
        assert np.array_equal(a.prev_pts, b.prev_pts) and np.array_equal(bits(a.next_pts), bits(b.next_pts))
        assert np.array_equal(a.status, b.status)
        assert (a.inlier_mask is None) == (b.inlier_mask is None)
        if a.inlier_mask is not None:
            assert np.array_equal(a.inlier_mask, b.inlier_mask)
            assert np.array_equal(a.affine, b.affine)
        assert (a.detected is None) == (b.detected is None)
        if a.detected is not None:
            assert np.array_equal(a.detected, b.detected)
    assert len(port.output_records) == len(ref.output_records)
    for a, b in zip(port.output_records, ref.output_records):
        assert (a.index, a.path_len, a.radius, a.intent) == (b.index, b.path_len, b.radius, b.intent)
        assert np.array_equal(bits(a.smoothed), bits(b.smoothed))
        assert (a.T is None) == (b.T is None)
        if a.T is not None:
            assert np.array_equal(bits(a.T), bits(b.T))
    for i, (a, b) in enumerate(zip(o_port, o_ref)):
        assert a.shape == b.shape and np.array_equal(a, b), f"output frame {i} differs"
    return port, ref


CASES = {
    "box_default": (1280, 720, 48, 11, Parameters()),
    "box_r10": (1280, 720, 40, 78, Parameters(smoothingRadius=10)),
    "gaussian": (1280, 720, 40, 78, Parameters(smoothingRadius=10, smoothingMethod="gaussian", gaussianSigma=2.0)),
    "gaussian_wide_sigma_falls_back": (640, 360, 30, 5, Parameters(smoothingRadius=8, smoothingMethod="gaussian", gaussianSigma=1.0)),
    "kalman_hlock": (1280, 720, 40, 79, Parameters(smoothingRadius=10, smoothingMethod="kalman", horizonLock=True)),
    "border_reflect": (1280, 720, 14, 77, Parameters(smoothingRadius=5, borderType="reflect", borderSize=24)),
    "border_wrap": (640, 360, 14, 77, Parameters(smoothingRadius=5, borderType="wrap", borderSize=10)),
    "border_replicate": (640, 360, 14, 77, Parameters(smoothingRadius=5, borderType="replicate", borderSize=7)),
    "border_reflect101": (640, 360, 14, 77, Parameters(smoothingRadius=5, borderType="reflect_101", borderSize=9)),
    "border_black": (640, 360, 14, 77, Parameters(smoothingRadius=5, borderType="black", borderSize=12)),
    "cropzoom": (1280, 720, 14, 3000, Parameters(smoothingRadius=5, cropNZoom=True, borderSize=30)),
    "cropzoom_reflect_forced_black": (640, 360, 14, 3000, Parameters(smoothingRadius=5, cropNZoom=True, borderSize=12, borderType="reflect")),
    "fade": (640, 360, 20, 91, Parameters(smoothingRadius=5, borderType="fade", borderSize=16, fadeDuration=6)),
    "drone": (1280, 720, 40, 17, Parameters(smoothingRadius=12, droneHighFreqMode=True, horizonLock=True)),
    "drone_no_hlock": (1280, 720, 30, 18, Parameters(smoothingRadius=30, droneHighFreqMode=True, hfShakePx=2.5)),
    "adaptive": (640, 360, 40, 91, Parameters(smoothingRadius=12, adaptiveSmoothing=True)),
    "corners_custom": (1280, 720, 24, 21, Parameters(smoothingRadius=6, maxCorners=60, qualityLevel=0.05, minDistance=12.0)),
    "1080p_r15": (1920, 1080, 24, 2000, Parameters(smoothingRadius=15)),
    "odd_size": (1000, 562, 20, 40, Parameters(smoothingRadius=7)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_restatement_equals_reference_cpp(name):
    w, h, n, seed, params = CASES[name]
    assert_same_run(synthclip.make_clip(w, h, n, seed), params)


def test_reference_degenerate_inputs():
    """flat frames (no corners -> zero transforms), a scene cut (LK loses everything), clean() and reuse."""
    w, h = 640, 360
    a = synthclip.make_clip(w, h, 10, 71)
    b = synthclip.make_clip(w, h, 8, 72)
    flat = np.full((6, h, w, 3), 90, np.uint8)
    clip = np.concatenate([a, flat, b])
    port, ref = assert_same_run(clip, Parameters(smoothingRadius=5))
    assert any(len(r.status) == 0 or not r.status.any() for r in port.frame_records)


def test_reference_intents_are_reached():
    """Clips built to leave NORMAL: a steady pan reaches DELIBERATE_PAN; an alternating roll about the frame origin with
    occasional translation jumps reaches SHAKE_REMOVAL and FOLLOW_ACTION — in the reference's own analyzeMotionIntent
    (Stabilizer.cpp:1676-1719) — and the restatement takes the same branches frame for frame."""
    _, ref = assert_same_run(synthclip.pan_clip(640, 360, 50, 123), Parameters(smoothingRadius=5))
    assert 1 in {r.intent for r in ref.output_records}, "DELIBERATE_PAN not reached"
    _, ref = assert_same_run(synthclip.shake_clip(640, 360, 60, 5), Parameters(smoothingRadius=5))
    intents = [r.intent for r in ref.output_records]
    assert intents.count(2) >= 20 and 3 in intents, f"SHAKE_REMOVAL / FOLLOW_ACTION not reached: {intents}"


def test_gaussian_short_path_is_undefined_in_the_reference_and_defined_here():
    """Quirk B-Q8: with gate-1 <= centre the reference's gaussianFilterConvolve reads outside the path vector
    (Stabilizer.cpp:1392-1401) — undefined behaviour.  The restatement (and the CUDA path) fall back to the box filter until
    the path is longer than the kernel centre.  Everything that does not depend on those out-of-bounds reads still agrees
    bit for bit: all transforms, and every output popped once the path is long enough."""
    import oracle
    params = Parameters(smoothingRadius=7, smoothingMethod="gaussian", gaussianSigma=3.5)      # centre 10, first pop at 6 samples
    assert oracle.gaussian_reads_out_of_bounds(params)
    assert not oracle.gaussian_reads_out_of_bounds(Parameters(smoothingRadius=10, smoothingMethod="gaussian", gaussianSigma=2.0))
    clip = synthclip.make_clip(640, 360, 40, 607)
    o_port, port = run_clip(clip, params)
    o_ref, ref = ref_lib.run_clip(clip, params)
    assert np.array_equal(bits(np.array(port.transforms).reshape(-1, 3)), bits(ref.transforms()))
    centre, checked = 10, 0
    for a, b, x, y in zip(port.output_records, ref.output_records, o_port, o_ref):
        if a.T is not None and a.path_len > centre + 1:
            assert np.array_equal(bits(a.smoothed), bits(b.smoothed)) and np.array_equal(x, y), f"output {a.index}"
            checked += 1
    assert checked >= 20


# ----------------------------------------------------------------------------- pure host functions, on their own
@pytest.fixture(scope="module")
def ref_box():
    return ref_lib.RefStabilizer(Parameters(smoothingRadius=30), record=False)


@pytest.mark.parametrize("radius,drone", [(1, False), (5, False), (8, False), (30, False), (3, True), (25, True), (80, True)])
def test_box_filter_convolve(radius, drone):
    ref = ref_lib.RefStabilizer(Parameters(smoothingRadius=radius, droneHighFreqMode=drone), record=False)
    rng = np.random.default_rng(radius)
    for n in (0, 1, 2, 7, 9, 11, 51, 400):
        path = np.cumsum(rng.normal(0, 3, n)).astype(np.float32)
        got = ref.box_filter(path)
        want = StabilizerRef._box(path, radius, drone) if n else np.zeros(0, np.float32)
        assert np.array_equal(bits(got), bits(want)), (radius, drone, n)
        for i in range(0, n, 5):
            assert bits(StabilizerRef.box_at(path, radius, i, drone)) == bits(want[i])


@pytest.mark.parametrize("sigma", [0.4, 1.0, 2.0, 3.7, 15.0])
def test_gaussian_filter_convolve(ref_box, sigma):
    rng = np.random.default_rng(int(sigma * 10))
    ksz = max(3, int(np.ceil(np.float32(6) * np.float32(sigma))))
    ksz += ksz % 2 == 0
    for n in (ksz // 2 + 1, ksz, 64, 333):            # n <= centre reads out of bounds in the reference (B-Q8): not driven
        path = np.cumsum(rng.normal(0, 3, n)).astype(np.float32)
        assert np.array_equal(bits(ref_box.gaussian_filter(path, sigma)), bits(StabilizerRef._gaussian(path, sigma))), (sigma, n)


def test_kalman_filter_smooth(ref_box):
    port = StabilizerRef(Parameters())
    rng = np.random.default_rng(3)
    for n in (1, 2, 17, 200):
        path = np.cumsum(rng.normal(0, 3, n)).astype(np.float32)
        assert np.array_equal(bits(ref_box.kalman_filter(path)), bits(port._kalman(path))), n


def test_adaptive_radius_variance_consistency(ref_box):
    port = StabilizerRef(Parameters(smoothingRadius=30))
    rng = np.random.default_rng(4)
    for trial in range(200):
        n = int(rng.integers(0, 60))
        sc = float(rng.choice([0.01, 0.3, 2.0, 20.0]))
        px, py = (np.cumsum(rng.normal(0, sc, n)).astype(np.float32) for _ in range(2))
        pa = np.cumsum(rng.normal(0, sc * 0.002, n)).astype(np.float32)
        assert ref_box.adaptive_radius(px, py, pa) == port._adaptive_radius(px, py, pa)
        v = np.abs(rng.normal(2, sc, n)).astype(np.float32)
        assert bits(ref_box.variance(v)) == bits(port._variance(list(v)))
        assert bits(ref_box.consistency(v)) == bits(port._consistency(list(v)))


def test_motion_intent_and_strength(ref_box):
    port = StabilizerRef(Parameters())
    rng = np.random.default_rng(5)
    seen = set()
    for trial in range(300):
        n = int(rng.integers(1, 60))
        kind = trial % 4
        if kind == 0:       # steady pan
            t = np.stack([rng.normal(8, 0.3, n), rng.normal(1, 0.2, n), rng.normal(0, 0.001, n)], 1)
        elif kind == 1:     # rotational shake, erratic tiny translation
            t = np.stack([rng.normal(0, 1, n) * (rng.random(n) < 0.3), rng.normal(0, 1, n) * (rng.random(n) < 0.3), rng.normal(0, 0.02, n)], 1)
        elif kind == 2:     # follow action
            ang = rng.uniform(-np.pi, np.pi, n)
            mag = rng.uniform(4, 12, n)
            t = np.stack([mag * np.cos(ang), mag * np.sin(ang), rng.normal(0, 0.002, n)], 1)
        else:
            t = np.stack([rng.normal(0, 3, n), rng.normal(0, 3, n), rng.normal(0, 0.004, n)], 1)
        t = t.astype(np.float32)
        ref_box.set_transforms(t)
        port.transforms = [r.copy() for r in t]
        for idx in {1, n // 2, n - 1, n} - {0}:
            m = t[min(idx, n - 1)]
            got = ref_box.motion_intent(m, idx)
            assert got == port._intent(m, idx), (trial, idx)
            seen.add(got)
            if got == 0:      # the strength is consumed only on the NORMAL branch (Stabilizer.cpp:881-884), where it is 0.7f
                assert bits(ref_box.stabilization_strength(got, m)) == bits(np.float32(0.7))
    assert seen == {0, 1, 2, 3}, seen
    ref_box.set_transforms(np.zeros((0, 3), np.float32))


def test_adapt_smoothing_radius():
    ref = ref_lib.RefStabilizer(Parameters(smoothingRadius=20, adaptiveSmoothing=True, minSmoothingRadius=4, maxSmoothingRadius=47), record=False)
    port = StabilizerRef(Parameters(smoothingRadius=20, adaptiveSmoothing=True, minSmoothingRadius=4, maxSmoothingRadius=47))
    rng = np.random.default_rng(6)
    for _ in range(200):
        m = np.array([rng.normal(0, 25), rng.normal(0, 25), 0], np.float32)
        port.transforms = [m, m, m]
        port._update_adaptive()
        assert ref.adapt_smoothing_radius(m) == port.p.smoothingRadius


def test_drone_chain_and_analysis_size():
    kw = dict(droneHighFreqMode=True, horizonLock=True, hfShakePx=1.5, hfDeadZoneThreshold=2.0, hfFreezeDuration=4)
    ref = ref_lib.RefStabilizer(Parameters(**kw), record=False)
    port = StabilizerRef(Parameters(**kw))
    rng = np.random.default_rng(7)
    for k in range(400):
        sc = 0.4 if (k // 40) % 2 == 0 else 4.0
        t = np.array([rng.normal(0, sc), rng.normal(0, sc), rng.normal(0, 0.004)], np.float32)
        assert np.array_equal(bits(ref.drone_chain(t)), bits(port._hf_filters(t))), k
    assert ref.drone_analysis_size(1920, 1080) == (960, 540)
    assert ref.drone_analysis_size(3840, 2160) == (960, 540)
    assert ref.drone_analysis_size(1280, 720) == (960, 540)


# ----------------------------------------------------------------------------- the committed goldens, from the reference
GOLDEN_CASES = {
    "cfg2_1080p_r15": dict(w=1920, h=1080, n=48, seed=2000, params=Parameters(smoothingRadius=15)),
    "border_reflect_720p": dict(w=1280, h=720, n=12, seed=77, params=Parameters(smoothingRadius=5, borderType="reflect", borderSize=24)),
    "gaussian_720p": dict(w=1280, h=720, n=40, seed=78, params=Parameters(smoothingRadius=10, smoothingMethod="gaussian", gaussianSigma=2.0)),
    "kalman_hlock_720p": dict(w=1280, h=720, n=40, seed=79, params=Parameters(smoothingRadius=10, smoothingMethod="kalman", horizonLock=True)),
    "cfg1_720p_default": dict(w=1280, h=720, n=300, seed=1234, params=Parameters()),
    "cfg3_4k_cropzoom": dict(w=3840, h=2160, n=10, seed=3000, params=Parameters(smoothingRadius=5, cropNZoom=True, borderSize=30)),
}


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_golden_fixture_is_what_the_reference_produces(name):
    c = GOLDEN_CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    clip = synthclip.make_clip(c["w"], c["h"], c["n"], c["seed"])
    outs, ref = ref_lib.run_clip(clip, c["params"])
    assert np.array_equal(g["first_corners"], ref.first_corners)
    assert np.array_equal(bits(g["transforms"]), bits(ref.transforms()))
    assert np.array_equal(bits(g["path"]), bits(ref.path()))
    for i, r in enumerate(ref.frame_records):
        n = len(r.status)
        assert g["prev_n"][i] == n and np.array_equal(g["status"][i, :n], r.status)
        assert np.array_equal(bits(g["next_pts"][i, :n]), bits(r.next_pts))
        if r.inlier_mask is None:
            assert g["inlier_n"][i] == -1
        else:
            m = len(r.inlier_mask)
            assert g["inlier_n"][i] == m and np.array_equal(g["inlier_mask"][i, :m], r.inlier_mask)
        if r.detected is None:
            assert g["detected_n"][i] == -1
        else:
            assert np.array_equal(g["detected"][i, :g["detected_n"][i]], r.detected)
    o = ref.output_records
    assert np.array_equal(g["out_index"], [r.index for r in o])
    assert np.array_equal(g["out_radius"], [r.radius for r in o])
    assert np.array_equal(g["out_intent"], [r.intent for r in o])
    assert np.array_equal(bits(g["out_smoothed"]), bits(np.stack([r.smoothed for r in o])))
    assert np.array_equal(bits(g["out_T"]), bits(np.stack([np.zeros((2, 3), np.float32) if r.T is None else r.T for r in o])))
    assert np.array_equal(g["out_crc"], [zlib.crc32(np.ascontiguousarray(f).tobytes()) & 0xFFFFFFFF for f in outs])


@pytest.mark.parametrize("kw", [dict(), dict(canvasScaleFactor=1.3, adaptiveCanvasSize=False),
                                dict(canvasScaleFactor=1.0, adaptiveCanvasSize=False, temporalBufferSize=4, edgeBlendRadius=7, canvasBlendWeight=0.45),
                                dict(minCanvasScale=1.0, maxCanvasScale=1.9)])
def test_virtual_canvas_restatement_equals_compiled_reference(kw):
    """oracle/virtual_canvas_ref.py against the reference's own applyVirtualCanvasStabilization (Stabilizer.cpp:2066-2443,
    compiled unmodified): the same frames and corrections give bit-equal frames, whole-canvas fills, dark-object fills and the
    adaptive canvas size included."""
    from oracle.virtual_canvas_ref import VirtualCanvasRef
    w, h, n = 320, 180, 12
    clip = np.maximum(synthclip.make_clip(w, h, n, 5), 6)
    clip[:, 60:75, 100:130] = 0
    clip[:, 0:9, 200:230] = 0
    clip[4:, 120:160, 20:60] = 1
    rng = np.random.default_rng(3)
    corr = np.stack([rng.uniform(-40, 40, n), rng.uniform(-25, 25, n), rng.uniform(-0.03, 0.03, n)], axis=1).astype(np.float32)
    corr[::5, :2] = np.round(corr[::5, :2])
    recent = (rng.uniform(-1, 1, (35, 3)) * 75).astype(np.float32)
    ref = ref_lib.RefStabilizer(dict(enableVirtualCanvas=True, **kw), record=False)
    ref.set_transforms(recent)
    mine = VirtualCanvasRef(**kw)
    fills = 0
    for k in range(n):
        want = ref.vc_apply(clip[k], corr[k])
        got = mine.apply(clip[k], corr[k], recent)
        assert np.array_equal(want, got), f"frame {k} differs"
        fills += len(mine.filled)
    assert mine.scale == ref.vc_last()[1]
    if kw.get("canvasScaleFactor", 1.5) < 1.4:
        assert fills > 0

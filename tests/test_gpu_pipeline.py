"""Whole-path parity through the reference-shaped API (vs::Stabilizer mirror over the C-ABI) against
the oracle run live and against the committed golden fixtures: bit-exact corner lists, LK status and
RANSAC inlier masks; transforms within 1e-3 px of corner displacement; frames within 1 LSB."""
import os
import zlib

import numpy as np

import synthclip
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HALF_DIAG = 1102.0      # half diagonal of the 960x540 analysis image: rad -> px of corner displacement


@pytest.fixture(scope="module")
def vsb():
    import __graft_entry__
    __graft_entry__.build()
    import video_stab_b200
    assert torch.cuda.is_available()
    return video_stab_b200


def _run(vsb, clip, params):
    st = vsb.Stabilizer(params)
    outs = []
    for f in clip:
        o = st.stabilize(f)
        if o is not None:
            outs.append(o)
    while True:
        o = st.flush()
        if o is None:
            break
        outs.append(o)
    return outs, st


def _crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def _check_against_golden(vsb, name, w, h, n, seed, params, frame_tol=1):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    clip = synthclip.make_clip(w, h, n, seed)
    assert np.array_equal(np.array([_crc(f) for f in clip], np.uint32), g["input_crc"]), \
        "synthetic clip differs from the one the goldens were made from"
    outs, st = _run(vsb, clip, params)
    nf, no = st.counts()
    assert nf == len(g["transforms"]) and no == len(g["out_index"]) == len(outs)
    assert np.array_equal(st.first_corners(), g["first_corners"])
    exact_pts = 0
    for i in range(nf):
        rec = st.frame_record(i)
        pts = st.frame_points(i)
        pn = int(g["prev_n"][i])
        assert rec.n_prev_pts == pn
        assert np.array_equal(pts["prev"], g["prev_pts"][i, :pn])
        assert np.array_equal(pts["status"], g["status"][i, :pn]), f"frame {i}: LK status"
        ok = pts["status"] == 1
        assert np.abs(pts["next"][ok] - g["next_pts"][i, :pn][ok]).max(initial=0) < 1e-3
        exact_pts += int(np.array_equal(pts["next"][ok].view(np.uint32), g["next_pts"][i, :pn][ok].view(np.uint32)))
        mn = int(g["inlier_n"][i])
        if mn >= 0:
            assert np.array_equal(pts["inlier_mask"], g["inlier_mask"][i, :mn]), f"frame {i}: inlier mask"
        else:
            assert pts["inlier_mask"] is None
        dn = int(g["detected_n"][i])
        if dn >= 0:
            assert np.array_equal(pts["detected"], g["detected"][i, :dn]), f"frame {i}: corner list"
        else:
            assert pts["detected"] is None
        d = np.abs(np.asarray(rec.transform) - g["transforms"][i])
        assert d[0] < 1e-3 and d[1] < 1e-3 and d[2] * HALF_DIAG < 1e-3, f"frame {i}: transform {d}"
        dp = np.abs(np.asarray(rec.path) - g["path"][i])
        assert dp[0] < 2e-3 and dp[1] < 2e-3 and dp[2] * HALF_DIAG < 2e-3
    worst = 0
    exact_frames = 0
    for k in range(no):
        r = st.output_record(k)
        assert r.index == g["out_index"][k] and r.passthrough == g["out_passthrough"][k]
        if not r.passthrough:
            assert r.radius == g["out_radius"][k] and r.intent == g["out_intent"][k]
            dT = np.abs(np.asarray(r.T).reshape(2, 3) - g["out_T"][k])
            assert dT[:, 2].max() < 1e-3 and dT[:, :2].max() * HALF_DIAG < 1e-3
        assert tuple(outs[k].shape) == tuple(g["out_shape"][k])
        row = outs[k][outs[k].shape[0] // 2, 100:260, :].astype(np.int16)
        worst = max(worst, int(np.abs(row - g["out_row"][k].astype(np.int16)).max()))
        exact_frames += int(_crc(outs[k]) == g["out_crc"][k])
    assert worst <= frame_tol
    return exact_pts / max(nf, 1), exact_frames / max(no, 1)


def test_config1_720p_300_frames_vs_golden(vsb):
    """BASELINE config 1: 1280x720, 300 frames, defaults."""
    lk_exact, frames_exact = _check_against_golden(vsb, "cfg1_720p_default", 1280, 720, 300, 1234, vsb.Parameters())
    print(f"cfg1: LK point sets bit-exact on {lk_exact:.1%} of frames, output frames bit-exact on {frames_exact:.1%}")
    assert lk_exact > 0.95


def test_config2_1080p_radius15_vs_golden(vsb):
    _check_against_golden(vsb, "cfg2_1080p_r15", 1920, 1080, 48, 2000, vsb.Parameters(smoothingRadius=15))


def test_config3_4k_cropzoom_vs_golden(vsb):
    _check_against_golden(vsb, "cfg3_4k_cropzoom", 3840, 2160, 10, 3000,
                          vsb.Parameters(smoothingRadius=5, cropNZoom=True, borderSize=30))


def test_border_reflect_vs_golden(vsb):
    _check_against_golden(vsb, "border_reflect_720p", 1280, 720, 12, 77,
                          vsb.Parameters(smoothingRadius=5, borderType="reflect", borderSize=24))


def test_gaussian_vs_golden(vsb):
    _check_against_golden(vsb, "gaussian_720p", 1280, 720, 40, 78,
                          vsb.Parameters(smoothingRadius=10, smoothingMethod="gaussian", gaussianSigma=2.0))


def test_kalman_horizon_lock_vs_golden(vsb):
    _check_against_golden(vsb, "kalman_hlock_720p", 1280, 720, 40, 79,
                          vsb.Parameters(smoothingRadius=10, smoothingMethod="kalman", horizonLock=True))


def test_live_oracle_full_frames(vsb, cv2_noopt):
    """Full-frame comparison against the oracle run on this box (not just the golden row slices)."""
    from oracle import run_clip
    from oracle.stabilizer_ref import Parameters
    clip = synthclip.make_clip(1280, 720, 24, 555)
    outs, st = _run(vsb, clip, vsb.Parameters(smoothingRadius=8))
    ref_outs, ref = run_clip(clip, Parameters(smoothingRadius=8))
    assert len(outs) == len(ref_outs) == 24
    band = 40       # border band: where the zero border blends in, a 1-ulp matrix difference is amplified
    exact = 0
    for k, (a, b) in enumerate(zip(outs, ref_outs)):
        assert a.shape == b.shape
        d = np.abs(a.astype(np.int16) - b.astype(np.int16))
        assert d[band:-band, band:-band].max() <= 1, f"output {k}: {d[band:-band, band:-band].max()} LSB"
        assert d.max() <= 12 and (d > 1).mean() < 1e-3
        exact += int(d.max() == 0)
    print(f"bit-exact output frames: {exact}/24")
    assert np.array_equal(outs[-1], clip[-1])        # last frame has no transform: passed through (:774-780)


def test_adaptive_smoothing_gate(vsb, cv2_noopt):
    from oracle import run_clip
    from oracle.stabilizer_ref import Parameters
    clip = synthclip.make_clip(640, 360, 40, 91)
    kw = dict(smoothingRadius=12, adaptiveSmoothing=True, minSmoothingRadius=6, maxSmoothingRadius=20)
    ref_outs, ref = run_clip(clip, Parameters(**kw), flush=False)
    st = vsb.Stabilizer(vsb.Parameters(**kw))
    produced = [st.stabilize(f) is not None for f in clip]
    ref_st = __import__("oracle.stabilizer_ref", fromlist=["StabilizerRef"]).StabilizerRef(Parameters(**kw))
    ref_produced = [ref_st.stabilize(f) is not None for f in clip]
    assert produced == ref_produced


def test_empty_frame_and_clean(vsb):
    st = vsb.Stabilizer(vsb.Parameters(smoothingRadius=5))
    assert st.stabilize(None) is None
    assert st.flush() is None
    clip = synthclip.make_clip(640, 360, 8, 3)
    a = [st.stabilize(f) for f in clip]
    st.clean()
    b = [st.stabilize(f) for f in clip]
    assert [x is None for x in a] == [x is None for x in b] == [True] * 4 + [False] * 4
    for x, y in zip(a, b):
        if x is not None:
            assert np.array_equal(x, y)             # clean() restores a fresh stabilizer


def test_device_api_and_batch_equal_single(vsb):
    """Config 4 in miniature: 3 streams in one lock-step batch == 3 independent stabilizers."""
    w, h, n, S = 640, 360, 14, 3
    clips = [synthclip.make_clip(w, h, n, 2000 + s) for s in range(S)]
    params = vsb.Parameters(smoothingRadius=6)
    singles = [_run(vsb, c, params)[0] for c in clips]
    d_clips = [torch.from_numpy(c).cuda() for c in clips]
    d_out = torch.zeros((S, n, h, w, 3), dtype=torch.uint8, device="cuda")
    batch = vsb.StabilizerBatch(params, S)
    k = 0
    for i in range(n):
        r = batch.push_device([d_clips[s][i].data_ptr() for s in range(S)], w, h, w * 3,
                              [d_out[s, k].data_ptr() for s in range(S)], w * 3, h * w * 3, borrow=True)
        if r is not None:
            k += 1
    while True:
        r = batch.flush_device([d_out[s, min(k, n - 1)].data_ptr() for s in range(S)], w * 3, h * w * 3)
        if r is None:
            break
        k += 1
    batch.sync()
    assert k == n
    got = d_out.cpu().numpy()
    for s in range(S):
        for i in range(n):
            assert np.array_equal(got[s, i], singles[s][i]), f"stream {s} frame {i}"
    assert batch.launch_count() > 0


@pytest.mark.parametrize("kw", [dict(smoothingRadius=6), dict(smoothingRadius=5, borderType="reflect", borderSize=12),
                                dict(smoothingRadius=7, cropNZoom=True, borderSize=16),
                                dict(smoothingRadius=5, adaptiveSmoothing=True),
                                dict(smoothingRadius=5, enableVirtualCanvas=True, borderSize=8),
                                dict(smoothingRadius=6, enableVirtualCanvas=True, canvasScaleFactor=1.25, adaptiveCanvasSize=False)])
def test_push_many_equals_per_frame_push(vsb, kw):
    """The pipelined host call (copy-in / compute / copy-out streams, staging rings) returns exactly the frames
    of n stabilize() calls + flush(), for ragged chunk sizes, page-locked and pageable buffers, more frames than the
    36-slot device ring, and the single-stream adaptive mode."""
    w, h, n = 640, 360, 90
    clip = synthclip.make_clip(w, h, n, 77)
    ref, st0 = _run(vsb, clip, vsb.Parameters(**kw))
    for pinned in (True, False):
        st = vsb.Stabilizer(vsb.Parameters(**kw))
        src = torch.from_numpy(clip).pin_memory().numpy() if pinned else clip
        got = []
        pos = 0
        for chunk in (1, 3, 40, 17, 29):
            got += [f.copy() for f in st.stabilize_many(src[pos:pos + chunk])]
            pos += chunk
        assert pos == n
        while True:
            tail = st.flush_many(8)
            if len(tail) == 0:
                break
            got += [f.copy() for f in tail]
        assert len(got) == len(ref) == n
        for i, (a, b) in enumerate(zip(got, ref)):
            assert a.shape == b.shape and np.array_equal(a, b), f"pinned={pinned}: frame {i} differs"
        for i in range(n - 1):
            assert list(st.frame_record(i).transform) == list(st0.frame_record(i).transform)


@pytest.mark.parametrize("w,h,b,dur", [(640, 360, 16, 30), (644, 362, 10, 4)])
def test_fade_border_vs_live_oracle(vsb, w, h, b, dur):
    """border_type "fade" (Stabilizer.cpp:914-978, 1070-1106): history blend before the warp, history update after
    it, whole-frame mask quirk included.  Bit-exact against the oracle (cv::addWeighted on its SIMD/FMA path);
    the last frame comes back un-warped and un-bordered."""
    from oracle import run_clip
    from oracle.stabilizer_ref import Parameters as RP
    n = 40
    clip = synthclip.make_clip(w, h, n, 91)
    kw = dict(smoothingRadius=5, borderType="fade", borderSize=b, fadeDuration=dur, fadeAlpha=0.25)
    ref, _ = run_clip(clip, RP(**kw))
    outs, st = _run(vsb, clip, vsb.Parameters(**kw))
    assert len(outs) == len(ref) == n
    # Frames are compared exactly, except for isolated pixels: the oracle takes cos/sin of the angle from glibc's
    # cosf/sinf, the device from the correctly rounded double function; when the two differ in the last bit the
    # fixed-point source coordinate of a handful of pixels moves by 1/32 px (transform tolerance 1e-3 px),
    # and with "fade" such a pixel then lingers in the history at +-1.
    total = bad = 0
    for i, (a, r) in enumerate(zip(outs, ref)):
        assert a.shape == r.shape, f"frame {i}: {a.shape} vs {r.shape}"
        d = np.abs(a.astype(np.int16) - r.astype(np.int16))
        total += d.size
        bad += int((d > 0).sum())
        assert int((d > 1).sum()) <= 24 and int(d.max()) <= 12, f"frame {i}: {int((d > 1).sum())} pixels off by > 1 LSB, max {int(d.max())}"
    assert bad <= total * 1e-5, f"{bad} of {total} bytes differ"
    # history survives clean() (the reference never resets borderHistory_): a second pass starts from the old history
    st.clean()
    again = [o for o in (st.stabilize(f) for f in clip[:12]) if o is not None]
    assert len(again) > 0 and not np.array_equal(again[0], outs[0])


def _calm_clip(vsb, w, h, n, seed):
    """A hovering-drone clip: sub-pixel to few-pixel vibration with calm stretches, so that every branch of the
    high-frequency chain (dead-zone entry / timed exit / motion exit, 1 % and 5 % shake residuals, median) runs."""
    rng = np.random.default_rng(seed)
    base = synthclip.base_texture(w, h, seed)
    amp = np.where((np.arange(n) // 12) % 2 == 0, 0.4, 2.5)
    poses = np.stack([np.cumsum(rng.normal(0, 0.15, n)) + rng.normal(0, 1, n) * amp,
                      np.cumsum(rng.normal(0, 0.15, n)) + rng.normal(0, 1, n) * amp,
                      rng.normal(0, 0.002, n)], axis=1)
    return np.stack([synthclip.render_frame(base, poses[k], w, h) for k in range(n)])


@pytest.mark.parametrize("case", ["shaky", "calm", "calm_lock"])
def test_drone_high_freq_mode_vs_live_oracle(vsb, cv2_noopt, case):
    """drone_high_freq_mode (Stabilizer.cpp:666-671, 2468-2529, 2605-2681, box radius clamp :1144-1146) at the
    960x540 analysis size: filtered transforms equal to the oracle's float for float, frames within 1 LSB."""
    from oracle import run_clip
    from oracle.stabilizer_ref import Parameters as RP
    n = 60
    if case == "shaky":
        clip = synthclip.make_clip(1280, 720, n, 17)
        kw = dict(smoothingRadius=12, droneHighFreqMode=True)
    else:
        clip = _calm_clip(vsb, 1280, 720, n, 23)
        kw = dict(smoothingRadius=30, droneHighFreqMode=True, hfFreezeDuration=4, horizonLock=(case == "calm_lock"))
    ref_outs, ref = run_clip(clip, RP(**kw))
    plain_outs, plain = run_clip(clip[:20], RP(**{**kw, "droneHighFreqMode": False}), flush=False)
    outs, st = _run(vsb, clip, vsb.Parameters(**kw))
    assert len(outs) == len(ref_outs) == n
    nf, no = st.counts()
    assert nf == len(ref.frame_records)
    changed = 0
    for i, fr in enumerate(ref.frame_records):
        rec = st.frame_record(i)
        d = np.abs(np.asarray(rec.transform, np.float32) - fr.transform)
        assert d[0] < 1e-3 and d[1] < 1e-3 and d[2] * HALF_DIAG < 1e-3, f"frame {i}: {rec.transform} vs {fr.transform}"
        if i < len(plain.frame_records):
            changed += int(not np.array_equal(fr.transform, plain.frame_records[i].transform))
    assert changed > 0, "the clip never exercised the high-frequency filters"
    for k, orec in enumerate(ref.output_records):
        r = st.output_record(k)
        assert r.index == orec.index and bool(r.passthrough) == (orec.T is None)
        if not r.passthrough:
            assert r.radius == orec.radius and r.intent == orec.intent
    band = 40
    for k, (a, b) in enumerate(zip(outs, ref_outs)):
        assert a.shape == b.shape
        d = np.abs(a.astype(np.int16) - b.astype(np.int16))
        assert d[band:-band, band:-band].max() <= 1, f"output {k}: {d[band:-band, band:-band].max()} LSB"
        assert d.max() <= 12 and (d > 1).mean() < 1e-3


@pytest.mark.parametrize("w,h,kw,asize", [(640, 360, dict(), (640, 360)), (640, 480, dict(), (640, 480)),
                                          (1920, 1200, dict(), (960, 600)), (1920, 1080, dict(hfAnalysisMaxWidth=1280), (1280, 720)),
                                          (1000, 562, dict(), (960, 538)), (1280, 720, dict(hfAnalysisMaxWidth=320), (320, 180)),
                                          (1280, 720, dict(hfAnalysisMaxWidth=648), (648, 364)), (644, 362, dict(), (644, 362)),
                                          (1280, 720, dict(hfAnalysisMaxWidth=650), (650, 364)), (1280, 720, dict(hfAnalysisMaxWidth=646), (646, 362))])
def test_drone_mode_other_analysis_sizes_vs_live_oracle(vsb, cv2_noopt, w, h, kw, asize):
    """calculateDroneAnalysisSize (Stabilizer.cpp:2447-2466): min(hf_analysis_max_width, width) wide, the frame's aspect ratio,
    both made even - pyramids, detector and tracker at that size instead of 960 x 540."""
    from oracle import run_clip
    from oracle.stabilizer_ref import Parameters as RP
    n = 36
    clip = synthclip.make_clip(w, h, n, 29)
    P = dict(smoothingRadius=8, droneHighFreqMode=True, **kw)
    ref_outs, ref = run_clip(clip, RP(**P))
    outs, st = _run(vsb, clip, vsb.Parameters(**P))
    assert len(outs) == len(ref_outs) == n
    half_diag = 0.5 * float(np.hypot(*asize))
    for i, fr in enumerate(ref.frame_records):
        rec = st.frame_record(i)
        if fr.prev_pts is not None and len(fr.prev_pts):
            assert rec.n_prev_pts == len(fr.prev_pts), f"frame {i}: {rec.n_prev_pts} points vs {len(fr.prev_pts)}"
            assert rec.n_tracked == int(np.count_nonzero(fr.status)), f"frame {i}: tracked {rec.n_tracked} vs {int(np.count_nonzero(fr.status))}"
        d = np.abs(np.asarray(rec.transform, np.float32) - fr.transform)
        assert d[0] < 1e-3 and d[1] < 1e-3 and d[2] * half_diag < 1e-3, f"frame {i}: {rec.transform} vs {fr.transform}"
    band = 40
    for k, (a, b) in enumerate(zip(outs, ref_outs)):
        assert a.shape == b.shape
        d = np.abs(a.astype(np.int16) - b.astype(np.int16))
        # transforms agree to 1e-3 px: a source coordinate may land in the neighbouring 1/32-px bin for isolated pixels
        inner = d[band:-band, band:-band]
        assert int((inner > 1).sum()) <= 24 and int(inner.max()) <= 12, f"output {k}: {int((inner > 1).sum())} pixels off by > 1 LSB, max {int(inner.max())}"
        assert (d > 1).mean() < 1e-3


def test_drone_mode_analysis_sizes_that_are_refused(vsb):
    st = vsb.Stabilizer(vsb.Parameters(droneHighFreqMode=True, hfAnalysisMaxWidth=48))
    with pytest.raises(vsb.VsError) as ei:
        st.stabilize(np.zeros((1080, 1920, 3), np.uint8))         # 48 x 26: OpenCV's tracker would drop pyramid levels
    assert ei.value.status == 7


@pytest.mark.parametrize("borrow,canvas", [(True, False), (False, False), (True, True), (False, True)])
def test_async_device_pipeline_is_deterministic(vsb, borrow, canvas):
    """The asynchronous device-pointer path keeps five streams busy (pyramid of frame n+1 while frame n is tracked,
    detections of two frames in flight, warp of an older frame).  Its frames must equal the per-frame synchronous host
    path - which serialises everything - on every repetition, with more frames than ring slots / event slots, and
    with the internal ring copy (borrow=False) as well as frames read in place."""
    w, h, n = 1280, 720, 100
    clip = synthclip.make_clip(w, h, n, 4321)
    # canvas: the virtual-canvas stage in its asynchronous form (the shift is read from the set-up block on the device)
    params = vsb.Parameters(smoothingRadius=7, enableVirtualCanvas=canvas)
    ref, _ = _run(vsb, clip, params)
    want = [zlib.crc32(o.tobytes()) for o in ref]
    d_clip = torch.from_numpy(clip).cuda()
    fb = h * w * 3
    for rep in range(3):
        d_out = torch.zeros((n, h, w, 3), dtype=torch.uint8, device="cuda")
        st = vsb.Stabilizer(params)
        k = 0
        for i in range(n):
            if st.push_device(d_clip[i].data_ptr(), w, h, w * 3, d_out[k].data_ptr(), w * 3, fb, borrow=borrow) is not None:
                k += 1
        while st.flush_device(d_out[min(k, n - 1)].data_ptr(), w * 3, fb) is not None:
            k += 1
        st.sync()
        assert k == n
        got = [zlib.crc32(f.tobytes()) for f in d_out.cpu().numpy()]
        assert got == want, f"rep {rep}: frames {[i for i in range(n) if got[i] != want[i]][:8]} differ"


def test_wait_event_orders_frames_produced_on_another_stream(vsb):
    """Stream-ordered hand-off (vs_stabilizer_wait_event): frames written by a slow producer on another CUDA stream
    are pushed without any host synchronisation; the handle's streams wait for the producer's event."""
    w, h, n = 640, 360, 24
    clip = synthclip.make_clip(w, h, n, 77)
    params = vsb.Parameters(smoothingRadius=5)
    ref, _ = _run(vsb, clip, params)
    d_clip = torch.from_numpy(clip).cuda()
    fb = h * w * 3
    prod = torch.cuda.Stream()
    ballast = torch.randn((4096, 4096), device="cuda")
    staged = torch.zeros((n, h, w, 3), dtype=torch.uint8, device="cuda")
    d_out = torch.zeros((n, h, w, 3), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    st = vsb.Stabilizer(params)
    k = 0
    for i in range(n):
        with torch.cuda.stream(prod):
            ballast = ballast @ ballast * 1e-4            # a few hundred microseconds ahead of the frame copy
            staged[i].copy_(d_clip[i], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(prod)
        st.wait_event(ev.cuda_event)
        if st.push_device(staged[i].data_ptr(), w, h, w * 3, d_out[k].data_ptr(), w * 3, fb, borrow=True) is not None:
            k += 1
    while st.flush_device(d_out[min(k, n - 1)].data_ptr(), w * 3, fb) is not None:
        k += 1
    st.sync()
    assert k == n == len(ref)
    got = d_out.cpu().numpy()
    for i in range(n):
        assert np.array_equal(got[i], ref[i]), f"frame {i}"


@pytest.mark.parametrize("borrow", [True, False])
def test_push_many_device_equals_per_frame_push(vsb, borrow):
    """vs_stabilizer_push_many_device is the per-frame loop moved inside the library: same frames, same records."""
    w, h, n = 640, 360, 50
    clip = synthclip.make_clip(w, h, n, 99)
    params = vsb.Parameters(smoothingRadius=6)
    ref, st0 = _run(vsb, clip, params)
    d_clip = torch.from_numpy(clip).cuda()
    fb = h * w * 3
    d_out = torch.zeros((n, h, w, 3), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    st = vsb.Stabilizer(params)
    k = st.push_many_device(d_clip.data_ptr(), fb, 20, w, h, w * 3, d_out.data_ptr(), w * 3, fb, borrow=borrow)
    k += st.push_many_device(d_clip[20].data_ptr(), fb, n - 20, w, h, w * 3, d_out[k].data_ptr(), w * 3, fb, borrow=borrow)
    while st.flush_device(d_out[min(k, n - 1)].data_ptr(), w * 3, fb) is not None:
        k += 1
    st.sync()
    assert k == n == len(ref)
    got = d_out.cpu().numpy()
    for i in range(n):
        assert np.array_equal(got[i], ref[i]), f"frame {i}"
    for i in range(st.counts()[0]):
        assert list(st.frame_record(i).transform) == list(st0.frame_record(i).transform)


@pytest.mark.parametrize("w,h,n,kw", [(1920, 1080, 150, dict(smoothingRadius=15)),
                                      (1280, 720, 120, dict(smoothingRadius=6, cropNZoom=True, borderSize=24))])
def test_multi_stream_engine_equals_single_stream(vsb, monkeypatch, w, h, n, kw):
    """The seven-stream engine (slot rings + one transitive guard per frame) against the same calls with every kernel
    serialised on one stream (VS_SINGLE_STREAM=1): frames and transforms must be identical, on every repetition."""
    clip = torch.from_numpy(synthclip.make_clip(w, h, 40, 31)).cuda()
    order = list(range(40)) + list(range(38, 0, -1))
    seq = clip[torch.tensor([order[k % len(order)] for k in range(n)], device="cuda")].contiguous()
    fb = w * h * 3
    torch.cuda.synchronize()
    want = None
    for rep in range(5):
        if rep == 0:
            monkeypatch.setenv("VS_SINGLE_STREAM", "1")
        else:
            monkeypatch.delenv("VS_SINGLE_STREAM", raising=False)
        if rep == 4:                                   # the motion step as one kernel (default: RANSAC half behind LK)
            monkeypatch.setenv("VS_SPLIT_MOTION", "0")
        out = torch.zeros((n, h, w, 3), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        st = vsb.Stabilizer(vsb.Parameters(**kw))
        k = st.push_many_device(seq.data_ptr(), fb, n, w, h, w * 3, out.data_ptr(), w * 3, fb, borrow=True)
        while k < n and st.flush_device(out[k].data_ptr(), w * 3, fb) is not None:
            k += 1
        st.sync()
        assert k == n
        got = [zlib.crc32(f.tobytes()) for f in out.cpu().numpy()]
        recs = [tuple(st.frame_record(i).transform) for i in range(st.counts()[0])]
        if want is None:
            want = (got, recs)
        else:
            assert got == want[0], f"rep {rep}: frames {[i for i in range(n) if got[i] != want[0][i]][:8]} differ"
            assert recs == want[1], f"rep {rep}: transforms differ"


def _sweep_case(seed):
    rng = np.random.default_rng(1000 + seed)
    w, h = [(640, 360), (644, 362), (960, 540), (1280, 720), (1000, 562)][int(rng.integers(0, 5))]
    kw = dict(smoothingRadius=int(rng.integers(5, 36)),
              smoothingMethod=["box", "box", "gaussian", "kalman"][int(rng.integers(0, 4))],
              gaussianSigma=float(rng.choice([1.0, 2.0, 3.5])),
              horizonLock=bool(rng.integers(0, 2)),
              maxCorners=int(rng.choice([40, 200, 600])),
              qualityLevel=float(rng.choice([0.01, 0.05])),
              minDistance=float(rng.choice([8.0, 15.0, 30.0])))
    border = int(rng.choice([0, 0, 9, 24]))
    if border:
        kw["borderSize"] = border
        mode = int(rng.integers(0, 6))
        if mode == 5:
            kw["cropNZoom"] = True
        else:
            kw["borderType"] = ["black", "reflect", "replicate", "wrap", "fade"][mode]
    # drone mode at any frame size: its analysis image is min(960, width) wide with the frame's aspect ratio (every third
    # 16:9 frame of at least 960 columns - the 960 x 540 analysis size - and every fourth of the others)
    big = (w * 9 == h * 16) and w >= 960
    if rng.integers(0, 3 if big else 4) == 0:
        kw["droneHighFreqMode"] = True
    return w, h, kw


@pytest.mark.parametrize("seed", range(int(os.environ.get("VS_SWEEP_SEEDS", "12"))))
def test_random_parameter_sweep_vs_live_oracle(vsb, cv2_noopt, seed):
    """Random points of the configuration space (smoother, radius, corner parameters, border mode, crop-zoom, fade,
    drone mode, frame sizes that do and do not take the aligned fast paths) against the oracle run live: corner lists
    and LK status bit-exact, transforms within 1e-3 px, frames within 1 LSB away from the border band."""
    from oracle import run_clip
    from oracle.stabilizer_ref import Parameters as RP
    w, h, kw = _sweep_case(seed)
    n = 44
    clip = synthclip.make_clip(w, h, n, 600 + seed)
    ref_outs, ref = run_clip(clip, RP(**kw))
    outs, st = _run(vsb, clip, vsb.Parameters(**kw))
    assert len(outs) == len(ref_outs) == n, kw
    nf, no = st.counts()
    assert nf == len(ref.frame_records)
    assert np.array_equal(st.first_corners(), ref.first_corners), kw
    for i, fr in enumerate(ref.frame_records):
        rec = st.frame_record(i)
        pts = st.frame_points(i)
        assert np.array_equal(pts["status"], fr.status), f"{kw} frame {i}: LK status"
        if fr.detected is not None:
            assert np.array_equal(pts["detected"], fr.detected), f"{kw} frame {i}: corner list"
        d = np.abs(np.asarray(rec.transform, np.float32) - fr.transform)
        assert d[0] < 1e-3 and d[1] < 1e-3 and d[2] * HALF_DIAG < 1e-3, f"{kw} frame {i}: {rec.transform} vs {fr.transform}"
    band = 48
    fade = kw.get("borderType") == "fade"
    for k, (a, b) in enumerate(zip(outs, ref_outs)):
        assert a.shape == b.shape, f"{kw} output {k}"
        d = np.abs(a.astype(np.int16) - b.astype(np.int16))
        inner = d[band:-band, band:-band]
        if fade:
            assert (inner > 1).mean() < 1e-4 and d.max() <= 12, f"{kw} output {k}"
        elif kw.get("droneHighFreqMode") and w < 960:
            # a small analysis image: the 1e-3 px transform tolerance moves isolated pixels into the next 1/32-px bin
            assert int((inner > 1).sum()) <= 24 and d.max() <= 12 and (d > 1).mean() < 1e-3, f"{kw} output {k}"
        else:
            assert inner.max() <= 1, f"{kw} output {k}: {inner.max()} LSB"
            assert d.max() <= 12 and (d > 1).mean() < 1e-3, f"{kw} output {k}"


def test_degenerate_inputs_vs_live_oracle(vsb, cv2_noopt):
    """Inputs that push the path through its fallbacks: flat frames (no corners: the no-key-points branch pushes a zero
    transform, Stabilizer.cpp:676-678), the return of texture (re-detection), a scene cut (tracking / RANSAC failure:
    identity transform, :646-659), a very dark stretch and heavy noise."""
    from oracle import run_clip
    from oracle.stabilizer_ref import Parameters as RP
    w, h = 960, 540
    a = synthclip.make_clip(w, h, 12, 71)
    b = synthclip.make_clip(w, h, 10, 72)                       # unrelated texture: scene cut
    rng = np.random.default_rng(9)
    flat = np.full((6, h, w, 3), 127, np.uint8)
    dark = (a[:6].astype(np.float32) * 0.04).astype(np.uint8)
    noisy = np.clip(b[:6].astype(np.int16) + rng.normal(0, 40, (6, h, w, 3)), 0, 255).astype(np.uint8)
    clip = np.concatenate([a, flat, a[::-1][:8], b, dark, noisy, b[:6]])
    kw = dict(smoothingRadius=6)
    ref_outs, ref = run_clip(clip, RP(**kw))
    outs, st = _run(vsb, clip, vsb.Parameters(**kw))
    n = len(clip)
    assert len(outs) == len(ref_outs) == n
    assert st.counts()[0] == len(ref.frame_records)
    zero_tr = ident = 0
    for i, fr in enumerate(ref.frame_records):
        rec = st.frame_record(i)
        pts = st.frame_points(i)
        assert rec.n_prev_pts == len(fr.prev_pts), f"frame {i}: {rec.n_prev_pts} key points vs {len(fr.prev_pts)}"
        assert np.array_equal(pts["status"], fr.status), f"frame {i}: LK status"
        if fr.detected is not None:
            assert np.array_equal(pts["detected"], fr.detected), f"frame {i}: corner list"
        if fr.inlier_mask is not None:
            assert np.array_equal(pts["inlier_mask"], fr.inlier_mask), f"frame {i}: inlier mask"
        else:
            assert pts["inlier_mask"] is None, f"frame {i}: the oracle found no model"
            ident += 1
        d = np.abs(np.asarray(rec.transform, np.float32) - fr.transform)
        assert d[0] < 1e-3 and d[1] < 1e-3 and d[2] * HALF_DIAG < 1e-3, f"frame {i}: {rec.transform} vs {fr.transform}"
        zero_tr += int(len(fr.prev_pts) == 0)
    assert zero_tr >= 2 and ident >= 1, f"the clip no longer reaches the fallbacks ({zero_tr} empty, {ident} identity)"
    for k, (x, y) in enumerate(zip(outs, ref_outs)):
        dd = np.abs(x.astype(np.int16) - y.astype(np.int16))
        assert dd[40:-40, 40:-40].max() <= 1 and (dd > 1).mean() < 1e-3, f"output {k}"


@pytest.mark.parametrize("w,h", [(320, 240), (100, 60), (1366, 768), (2560, 1440), (1922, 1082), (3840, 2160)])
def test_unusual_frame_sizes_vs_live_oracle(vsb, cv2_noopt, w, h):
    """4:3, tiny, not-a-multiple-of-4, 1440p and 4K frames: the analysis image is always 960x540 (up- or down-scaled with
    cv::resize's arithmetic), the output stage takes the aligned TMA kernels or the fallbacks depending on the row pitch."""
    from oracle import run_clip
    from oracle.stabilizer_ref import Parameters as RP
    n = 16
    clip = synthclip.make_clip(w, h, n, 4000 + w)
    kw = dict(smoothingRadius=5, borderSize=6 if w < 2000 else 0, cropNZoom=(w == 1366))
    ref_outs, ref = run_clip(clip, RP(**kw))
    outs, st = _run(vsb, clip, vsb.Parameters(**kw))
    assert len(outs) == len(ref_outs) == n
    assert np.array_equal(st.first_corners(), ref.first_corners)
    for i, fr in enumerate(ref.frame_records):
        rec = st.frame_record(i)
        pts = st.frame_points(i)
        assert np.array_equal(pts["status"], fr.status), f"frame {i}: LK status"
        if fr.detected is not None:
            assert np.array_equal(pts["detected"], fr.detected), f"frame {i}: corner list"
        d = np.abs(np.asarray(rec.transform, np.float32) - fr.transform)
        assert d[0] < 1e-3 and d[1] < 1e-3 and d[2] * HALF_DIAG < 1e-3, f"frame {i}: {rec.transform} vs {fr.transform}"
    band = max(8, min(40, h // 8))
    for k, (a, b) in enumerate(zip(outs, ref_outs)):
        assert a.shape == b.shape, f"output {k}: {a.shape} vs {b.shape}"
        d = np.abs(a.astype(np.int16) - b.astype(np.int16))
        assert d[band:-band, band:-band].max() <= 1, f"output {k}: {d[band:-band, band:-band].max()} LSB"
        assert (d > 1).mean() < 2e-3, f"output {k}"


def test_adaptive_smoothing_frames_vs_live_oracle(vsb, cv2_noopt):
    """adaptive_smoothing (Stabilizer.cpp:691-693, 1461-1492, 1562-1574): the radius follows the last motion, moves the
    latency gate, and the run stays on one stream with one int read back per frame - frames and records as the oracle's."""
    from oracle import run_clip
    from oracle.stabilizer_ref import Parameters as RP
    w, h, n = 960, 540, 60
    clip = synthclip.make_clip(w, h, n, 91)
    kw = dict(smoothingRadius=12, adaptiveSmoothing=True, minSmoothingRadius=6, maxSmoothingRadius=20)
    ref_outs, ref = run_clip(clip, RP(**kw))
    outs, st = _run(vsb, clip, vsb.Parameters(**kw))
    assert len(outs) == len(ref_outs)
    assert st.counts()[0] == len(ref.frame_records)
    for i, fr in enumerate(ref.frame_records):
        d = np.abs(np.asarray(st.frame_record(i).transform, np.float32) - fr.transform)
        assert d[0] < 1e-3 and d[1] < 1e-3 and d[2] * HALF_DIAG < 1e-3, f"frame {i}"
    for k, orec in enumerate(ref.output_records):
        r = st.output_record(k)
        assert r.index == orec.index and bool(r.passthrough) == (orec.T is None), f"output {k}"
        if orec.T is not None:
            assert r.radius == orec.radius and r.intent == orec.intent, f"output {k}: radius {r.radius} vs {orec.radius}"
    for k, (a, b) in enumerate(zip(outs, ref_outs)):
        assert a.shape == b.shape
        d = np.abs(a.astype(np.int16) - b.astype(np.int16))
        assert d[40:-40, 40:-40].max() <= 1 and (d > 1).mean() < 1e-3, f"output {k}"


# ------------------------------------------------------------------------ BASELINE configs at their stated sizes
def test_oracle_on_this_box_is_the_compiled_reference():
    """The live-oracle tests of this module compare against the REFERENCE'S OWN Stabilizer.cpp (oracle/_ref, built in
    the build container by oracle/build_ref.py; the .so travels with the snapshot).  If it did not travel they fall
    back to the bit-identical Python restatement (tests/test_ref_pin.py); this test says which one ran."""
    import oracle
    if not oracle.reference_available():
        pytest.skip("oracle/_ref not present on this box: live-oracle tests used the Python restatement")
    assert oracle.oracle_kind() == "reference"


def test_config2_1080p_r15_live_oracle_full_frames(vsb, cv2_noopt):
    """BASELINE config 2 (the bench workload): 1920x1080, smoothing radius 15, 150 frames, EVERY output pixel of every
    frame against the oracle run live on this box.  Tolerances are the north star's: corner lists / status / inlier
    masks bit-exact, transforms within 1e-3 px of corner displacement, frames within 1 LSB outside the border band —
    and the fraction of bit-exact frames is asserted, not just reported."""
    from oracle import run_clip
    from oracle.stabilizer_ref import Parameters
    w, h, n = 1920, 1080, 150
    clip = synthclip.make_clip(w, h, n, 2000)
    outs, st = _run(vsb, clip, vsb.Parameters(smoothingRadius=15))
    ref_outs, ref = run_clip(clip, Parameters(smoothingRadius=15))
    assert len(outs) == len(ref_outs) == n
    assert np.array_equal(st.first_corners(), ref.first_corners)
    for i, r in enumerate(ref.frame_records):
        rec, pts = st.frame_record(i), st.frame_points(i)
        assert np.array_equal(pts["status"], r.status), f"frame {i}: LK status"
        if r.inlier_mask is not None:
            assert np.array_equal(pts["inlier_mask"], r.inlier_mask), f"frame {i}: inlier mask"
        if r.detected is not None:
            assert np.array_equal(pts["detected"], r.detected), f"frame {i}: corner list"
        d = np.abs(np.asarray(rec.transform) - r.transform)
        assert d[0] < 1e-3 and d[1] < 1e-3 and d[2] * HALF_DIAG < 1e-3, f"frame {i}: transform {d}"
    for k, r in enumerate(ref.output_records):
        o = st.output_record(k)
        assert o.index == r.index and bool(o.passthrough) == (r.T is None)
        if r.T is not None:
            assert (o.radius, o.intent) == (r.radius, r.intent), f"output {k}"
    band, exact, worst = 40, 0, 0
    for k, (a, b) in enumerate(zip(outs, ref_outs)):
        assert a.shape == b.shape
        d = np.abs(a.astype(np.int16) - b.astype(np.int16))
        inner = int(d[band:-band, band:-band].max())
        worst = max(worst, inner)
        assert inner <= 1, f"output {k}: {inner} LSB inside the border band"
        assert (d > 1).mean() < 1e-3
        exact += int(d.max() == 0)
    print(f"config 2: bit-exact output frames {exact}/{n}, worst interior difference {worst} LSB")
    assert exact >= int(0.98 * n), f"only {exact}/{n} output frames are bit-exact"


def test_config4_64_streams_batch_equals_singles(vsb):
    """BASELINE config 4 at its stated width: 64 concurrent 1080p streams in one lock-step batch == 64 independent
    stabilizers, every output frame compared on the device (40 frames per stream, radius 15).  The 64 streams are
    8 seeded clips x 8 start offsets (distinct content and phase per stream)."""
    w, h, n, S = 1920, 1080, 40, 64
    fb = h * w * 3
    params = vsb.Parameters(smoothingRadius=15)
    bases = [torch.from_numpy(synthclip.make_clip(w, h, n + 8, 2000 + s)).cuda() for s in range(8)]
    lane_clip = [bases[s % 8][s // 8: s // 8 + n] for s in range(S)]
    out_b = torch.zeros((S, n, h, w, 3), dtype=torch.uint8, device="cuda")
    batch = vsb.StabilizerBatch(params, S)
    k = 0
    for i in range(n):
        r = batch.push_device([lane_clip[s][i].data_ptr() for s in range(S)], w, h, w * 3,
                              [out_b[s, k].data_ptr() for s in range(S)], w * 3, fb, borrow=True)
        k += r is not None
    while True:
        r = batch.flush_device([out_b[s, min(k, n - 1)].data_ptr() for s in range(S)], w * 3, fb)
        if r is None:
            break
        k += 1
    batch.sync()
    assert k == n
    out_s = torch.zeros((n, h, w, 3), dtype=torch.uint8, device="cuda")
    for s in range(S):
        st = vsb.Stabilizer(params)
        m = st.push_many_device(lane_clip[s].data_ptr(), fb, n, w, h, w * 3, out_s.data_ptr(), w * 3, fb, borrow=True)
        while st.flush_device(out_s[min(m, n - 1)].data_ptr(), w * 3, fb) is not None:
            m += 1
        st.sync()
        assert m == n
        same = (out_b[s] == out_s).reshape(n, -1).all(1)
        assert bool(same.all()), f"stream {s}: frames {(~same).nonzero().flatten().tolist()[:5]} differ"
        for i in (0, n // 2, n - 2):
            assert list(batch.frame_record(s, i).transform) == list(st.frame_record(i).transform)
        del st


@pytest.mark.parametrize("kind", ["pan", "shake"])
def test_motion_intents_vs_live_oracle(vsb, cv2_noopt, kind):
    """Clips that leave MotionIntent::NORMAL (Stabilizer.cpp:1676-1719): a steady pan reaches DELIBERATE_PAN, an alternating
    roll about the frame origin with translation jumps reaches SHAKE_REMOVAL and FOLLOW_ACTION.  The device takes the same
    branch as the oracle on every output."""
    from oracle import run_clip
    from oracle.stabilizer_ref import Parameters
    w, h = 640, 360
    clip = synthclip.pan_clip(w, h, 50, 123) if kind == "pan" else synthclip.shake_clip(w, h, 60, 5)
    outs, st = _run(vsb, clip, vsb.Parameters(smoothingRadius=5))
    ref_outs, ref = run_clip(clip, Parameters(smoothingRadius=5))
    intents = [r.intent for r in ref.output_records]
    if kind == "pan":
        assert 1 in intents
    else:
        assert intents.count(2) >= 20 and 3 in intents
    for k, r in enumerate(ref.output_records):
        o = st.output_record(k)
        if r.T is not None:
            assert o.intent == r.intent and o.radius == r.radius, f"output {k}: intent {o.intent} vs {r.intent}"
            dT = np.abs(np.asarray(o.T).reshape(2, 3) - r.T)
            assert dT[:, 2].max() < 1e-3 and dT[:, :2].max() * HALF_DIAG < 1e-3
    for i, r in enumerate(ref.frame_records):
        assert np.array_equal(st.frame_points(i)["status"], r.status)
    for k, (a, b) in enumerate(zip(outs, ref_outs)):
        d = np.abs(a.astype(np.int16) - b.astype(np.int16))
        assert d[40:-40, 40:-40].max() <= 1, f"output {k}"


@pytest.mark.parametrize("kw", [dict(smoothingRadius=6), dict(smoothingRadius=5, borderType="reflect", borderSize=12)])
def test_too_small_output_buffer_is_recoverable(vsb, kw):
    """A too-small output buffer is reported BEFORE anything is queued or launched (VS_ERR_BUFFER_TOO_SMALL = 5): repeating
    the call with a large enough buffer continues the stream as if nothing had happened (same frames as a clean run)."""
    import ctypes as C
    from video_stab_b200._capi import lib
    w, h, n = 640, 360, 20
    clip = synthclip.make_clip(w, h, n, 12)
    ref, st0 = _run(vsb, clip, vsb.Parameters(**kw))
    st = vsb.Stabilizer(vsb.Parameters(**kw))
    b = kw.get("borderSize", 0)
    cap = (w + 2 * b) * (h + 2 * b) * 3
    big, small = np.empty(cap, np.uint8), np.empty(cap - 1, np.uint8)
    ow, oh, produced = C.c_int(), C.c_int(), C.c_int()
    got, failures = [], 0
    for i, f in enumerate(clip):
        f = np.ascontiguousarray(f)
        if i % 3 == 0:      # first try with a buffer one byte short
            rc = lib.vs_stabilizer_push(st._h, f.ctypes.data, w, h, w * 3, small.ctypes.data, 0, small.size, C.byref(ow), C.byref(oh), C.byref(produced))
            if rc != 0:
                assert rc == 5 and not produced.value
                failures += 1
            elif produced.value:
                raise AssertionError("a frame fitted into a too-small buffer")
            else:
                continue    # no output was due: the frame was accepted
        rc = lib.vs_stabilizer_push(st._h, f.ctypes.data, w, h, w * 3, big.ctypes.data, 0, big.size, C.byref(ow), C.byref(oh), C.byref(produced))
        assert rc == 0
        if produced.value:
            got.append(big[: ow.value * oh.value * 3].reshape(oh.value, ow.value, 3).copy())
    while True:
        rc = lib.vs_stabilizer_flush(st._h, small.ctypes.data, 0, w * h * 3 - 1, C.byref(ow), C.byref(oh), C.byref(produced))
        if rc == 0 and not produced.value:
            break
        assert rc == 5, "flush into a too-small buffer must be refused"
        rc = lib.vs_stabilizer_flush(st._h, big.ctypes.data, 0, big.size, C.byref(ow), C.byref(oh), C.byref(produced))
        assert rc == 0 and produced.value
        got.append(big[: ow.value * oh.value * 3].reshape(oh.value, ow.value, 3).copy())
    assert failures >= 3
    assert len(got) == len(ref) == n
    for i, (a, c) in enumerate(zip(got, ref)):
        assert a.shape == c.shape and np.array_equal(a, c), f"frame {i} differs after a refused call"
    for i in range(n - 1):
        assert list(st.frame_record(i).transform) == list(st0.frame_record(i).transform)


def test_lone_first_frame_flushed_from_the_device_ring(vsb):
    """push_device (copy mode) of ONE frame, then flush: the pass-through read on the public stream must be ordered behind
    the ring copy on the pyramid stream (there is no analysis chain between them for a lone frame).  Repeated to give a
    missing dependency a chance to show."""
    w, h = 1920, 1080
    fb = h * w * 3
    for rep in range(20):
        frame = torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, device="cuda")
        out = torch.zeros_like(frame)
        torch.cuda.synchronize()
        st = vsb.Stabilizer(vsb.Parameters(smoothingRadius=5))
        assert st.push_device(frame.data_ptr(), w, h, w * 3, out.data_ptr(), w * 3, fb, borrow=False) is None
        assert st.flush_device(out.data_ptr(), w * 3, fb) == (w, h)
        st.sync()
        assert torch.equal(out, frame), f"repetition {rep}"
        del st


def test_handles_on_two_devices_in_one_process(vsb):
    """Per-device kernel attributes (k_select / k_motion opt in to > 48 KB of dynamic shared memory, and that opt-in is per
    device): handles on cuda:0 and cuda:1 in one process produce the same frames as each other."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    w, h, n = 640, 360, 16
    clip = synthclip.make_clip(w, h, n, 808)
    outs = []
    for dev in (0, 1, 0):
        st = vsb.Stabilizer(vsb.Parameters(smoothingRadius=6), device=dev)
        got = [o for o in (st.stabilize(f) for f in clip) if o is not None]
        while True:
            o = st.flush()
            if o is None:
                break
            got.append(o)
        assert len(got) == n
        outs.append(got)
    for a, b, c in zip(*outs):
        assert np.array_equal(a, b) and np.array_equal(a, c)


@pytest.mark.parametrize("block", [5, 2])
def test_block_size_of_the_first_frame_detection_vs_live_oracle(vsb, cv2_noopt, block):
    """params.blockSize reaches only the first-frame goodFeaturesToTrack (Stabilizer.cpp:355-357); re-detections hard-code 3."""
    from oracle import run_clip
    from oracle.stabilizer_ref import Parameters
    clip = synthclip.make_clip(1280, 720, 16, 990 + block)
    kw = dict(smoothingRadius=5, blockSize=block, minDistance=12.0)
    outs, st = _run(vsb, clip, vsb.Parameters(**kw))
    ref_outs, ref = run_clip(clip, Parameters(**kw))
    assert np.array_equal(st.first_corners(), ref.first_corners)
    for i, r in enumerate(ref.frame_records):
        pts = st.frame_points(i)
        assert np.array_equal(pts["status"], r.status)
        if r.detected is not None:
            assert np.array_equal(pts["detected"], r.detected)
    for k, (a, b) in enumerate(zip(outs, ref_outs)):
        assert np.abs(a.astype(np.int16) - b.astype(np.int16))[40:-40, 40:-40].max() <= 1, f"output {k}"

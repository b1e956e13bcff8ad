"""vs::RollCorrection on the device against the REFERENCE'S OWN RollCorrection.cpp (oracle/_ref/libstages_ref.so,
cv::cuda:: calls served by the CPU functions of the same OpenCV — see oracle/mini_cv/opencv2/mini_cv_cuda.hpp).  Per frame: the small
gray image, the Canny edge map and the Hough line list (rho, theta, votes, order) are bit-exact; the smoothed angle agrees to 1e-9
degrees (double cos/sin of the roll angle are libdevice's on the device, glibc's in the oracle); the rotated frame is within 1 LSB and
bit-exact on all but a vanishing fraction of pixels."""
import numpy as np
import pytest

import synthclip

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def vsb():
    import __graft_entry__
    __graft_entry__.build()
    import video_stab_b200
    assert torch.cuda.is_available()
    return video_stab_b200


def _ref(params_kw):
    from oracle import ref_stages
    if not ref_stages.available():
        pytest.skip("oracle/_ref/libstages_ref.so not present")
    return ref_stages.RefRollCorrection(ref_stages.RollParameters(**params_kw))


CASES = {
    "1080p_default": (1920, 1080, 14, 7, {}),
    "1080p_config_yaml": (1920, 1080, 14, 8, dict(angleFilterMin=-70.0, angleFilterMax=70.0, angleDecay=0.98)),
    "4k": (3840, 2160, 6, 9, dict(angleFilterMin=-70.0, angleFilterMax=70.0)),
    "720p_low_threshold": (1280, 720, 12, 10, dict(houghThreshold=60, cannyThresholdLow=20.0, cannyThresholdHigh=60.0, angleFilterMin=-45.0, angleFilterMax=45.0)),
    "odd_size_half_scale": (1001, 563, 8, 11, dict(scaleFactor=0.5, houghThreshold=120, maxAngleChangeDeg=0.0)),
    "band_reaches_minus_90": (1280, 720, 8, 12, dict(angleFilterMin=-90.0, angleFilterMax=90.0, maxAngleChangeDeg=2.0)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_roll_correction_vs_reference(vsb, name):
    w, h, n, seed, kw = CASES[name]
    clip = synthclip.horizon_clip(w, h, n, seed)
    ref = _ref(kw)
    rc = vsb.RollCorrection(vsb.RollParameters(**kw))
    exact = 0
    for i, f in enumerate(clip):
        want = ref.correct(f)
        got = rc.autoCorrectRoll(f)
        dbg = rc.debug()
        cn, hg = ref.last("canny"), ref.last("hough")
        assert np.array_equal(dbg["gray"], cn["gray"]), f"frame {i}: small gray image"
        assert np.array_equal(dbg["edges"], cn["edges"]), f"frame {i}: Canny edges ({int((dbg['edges'] != cn['edges']).sum())} pixels differ)"
        assert len(dbg["lines"]) == len(hg["lines"]), f"frame {i}: {len(dbg['lines'])} lines vs {len(hg['lines'])}"
        assert np.array_equal(dbg["lines"].view(np.uint32), hg["lines"].view(np.uint32)), f"frame {i}: Hough lines"
        assert np.array_equal(dbg["votes"], hg["votes"]), f"frame {i}: Hough votes"
        assert abs(rc.state()["angle"] - ref.smoothed_angle) < 1e-9, f"frame {i}: angle {rc.state()['angle']} vs {ref.smoothed_angle}"
        d = np.abs(got.astype(np.int16) - want.astype(np.int16))
        assert d.max() <= 1 and (d > 0).mean() < 1e-4, f"frame {i}: max {d.max()}, {int((d > 0).sum())} pixels differ"
        exact += int(d.max() == 0)
    assert exact >= n - 1, f"only {exact}/{n} rotated frames are bit-exact"
    st = rc.state()
    assert st["launches"] == 9 * n


def test_roll_no_lines_decays_and_first_frame_resets(vsb):
    """flat frames: no edges, no lines -> the angle decays (RollCorrection.cpp:75-90); reset() restores sFirstFrame."""
    w, h = 1280, 720
    clip = synthclip.horizon_clip(w, h, 6, 3)
    flat = np.full((h, w, 3), 77, np.uint8)
    kw = dict(angleFilterMin=-70.0, angleFilterMax=70.0)
    ref = _ref(kw)
    rc = vsb.RollCorrection(vsb.RollParameters(**kw))
    for f in list(clip) + [flat] * 5:
        want, got = ref.correct(f), rc.autoCorrectRoll(f)
        assert abs(rc.state()["angle"] - ref.smoothed_angle) < 1e-9
        assert np.abs(got.astype(np.int16) - want.astype(np.int16)).max() <= 1
    assert rc.state()["n_lines"] == 0 and rc.state()["n_edges"] == 0
    a = rc.state()["angle"]
    assert a != 0.0
    rc.reset()
    rc.autoCorrectRoll(flat)
    assert rc.state()["angle"] == 0.0


def test_roll_device_api_matches_host_api(vsb):
    w, h = 1920, 1080
    clip = synthclip.horizon_clip(w, h, 5, 21)
    a, b = vsb.RollCorrection(), vsb.RollCorrection()
    s = torch.cuda.Stream()
    for f in clip:
        want = a.autoCorrectRoll(f)
        d_in = torch.from_numpy(f).cuda()
        d_out = torch.empty_like(d_in)
        torch.cuda.synchronize()
        b.correct_device(d_in.data_ptr(), w, h, w * 3, d_out.data_ptr(), w * 3, s.cuda_stream)
        s.synchronize()
        assert np.array_equal(d_out.cpu().numpy(), want)

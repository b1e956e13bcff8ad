"""The virtual-canvas output stage (`enable_virtual_canvas`, reference src/Stabilizer.cpp:1129-1134, 2066-2443) against the
reference's own code: the stage on its own, fed the same frames and corrections, must be bit-exact with the compiled
`applyVirtualCanvasStabilization`; inside the stabilizer the frames follow the corrections the device computes."""
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


def _ref_lib():
    from oracle import ref_lib
    if not ref_lib.available():
        pytest.skip("compiled reference (oracle/_ref) not present")
    return ref_lib


def _bright_clip(w, h, n, seed):
    """textured frames without any pixel of gray <= 1 (dark regions are painted explicitly)"""
    return np.maximum(synthclip.make_clip(w, h, n, seed), 6)


def _corrections(n, seed, amp, rot=0.02):
    rng = np.random.default_rng(seed)
    t = np.zeros((n, 3), np.float32)
    t[:, 0] = rng.uniform(-amp, amp, n)
    t[:, 1] = rng.uniform(-amp * 0.6, amp * 0.6, n)
    t[:, 2] = rng.uniform(-rot, rot, n)
    t[::7, :2] = np.round(t[::7, :2])                 # integer corrections: the int() casts sit on their boundaries
    t[3] = 0
    return t


def _run_stage(vsb, ref_lib, clip, corr, kw, recent=None, stride_pad=0):
    ref = ref_lib.RefStabilizer(dict(enableVirtualCanvas=True, **kw), record=False)
    if recent is not None:
        ref.set_transforms(recent)
    vc = vsb.VirtualCanvas(vsb.Parameters(enableVirtualCanvas=True, **kw))
    n, h, w = clip.shape[:3]
    pitch = w * 3 + stride_pad
    d_in = torch.zeros((h, pitch), dtype=torch.uint8, device="cuda")
    d_out = torch.zeros((h, pitch), dtype=torch.uint8, device="cuda")
    filled = []
    for k in range(n):
        want = ref.vc_apply(clip[k], corr[k])
        d_in[:, : w * 3] = torch.from_numpy(clip[k].reshape(h, w * 3)).cuda()
        vc.apply_device(d_in.data_ptr(), w, h, pitch, corr[k], d_out.data_ptr(), pitch, recent=recent)
        torch.cuda.synchronize()
        got = d_out[:, : w * 3].cpu().numpy().reshape(h, w, 3)
        filled.append(vc.info()["regions_filled"])
        assert want.shape == got.shape
        if not np.array_equal(want, got):
            d = (want != got).any(axis=2)
            ys, xs = np.nonzero(d)
            raise AssertionError(f"frame {k}: {int(d.sum())} pixels differ (rows {ys.min()}..{ys.max()}, cols {xs.min()}..{xs.max()}), "
                                 f"max |d| {int(np.abs(want.astype(int) - got).max())}, regions filled {filled[-1]}, T {corr[k]}")
    _, ref_scale, _ = ref.vc_last()
    assert abs(vc.info()["scale"] - ref_scale) == 0
    return filled, ref_scale


def _spots(clip):
    clip[:, 60:75, 100:130] = 0
    clip[:, 0:9, 200:230] = 0                        # touches the frame's edge
    clip[5:, 120:160, 20:60] = 1                     # gray 1 is still "empty"
    clip[8:, 100:104, 250:270] = 0                   # area 80: ignored
    return clip


def test_stage_default_parameters(vsb):
    """canvas 1.5x: the black surround encloses the frame, so its outer border is the only external contour and the one region
    is the whole canvas - which an older frame never covers by half: nothing is filled, the frame moves by whole pixels"""
    ref_lib = _ref_lib()
    w, h, n = 320, 180, 14
    filled, scale = _run_stage(vsb, ref_lib, _spots(_bright_clip(w, h, n, 3)), _corrections(n, 1, 30.0), dict())
    assert scale == 1.5 and max(filled) == 0


def test_stage_canvas_of_frame_size_fills_dark_objects(vsb):
    """canvas 1.0x: no surround, so dark objects inside the frame are external contours: each bounding rectangle is filled from
    the most recent older frame that covers it, with the edge ramp (mask to the host, border following there)"""
    ref_lib = _ref_lib()
    w, h, n = 320, 180, 14
    kw = dict(canvasScaleFactor=1.0, adaptiveCanvasSize=False)
    filled, scale = _run_stage(vsb, ref_lib, _spots(_bright_clip(w, h, n, 3)), _corrections(n, 1, 30.0), kw)
    assert scale == 1.0 and filled[0] == 0 and max(filled) == 3


def test_stage_partial_surround(vsb):
    """402 x 120 canvas around a 400 x 120 frame: two one-pixel columns of surround, not connected to each other"""
    ref_lib = _ref_lib()
    w, h, n = 400, 120, 8
    clip = _bright_clip(w, h, n, 11)
    clip[:, 30:50, 0:12] = 0                         # joins the left column
    clip[:, 70:90, 200:230] = 0
    kw = dict(canvasScaleFactor=1.005, adaptiveCanvasSize=False, edgeBlendRadius=3)
    filled, _ = _run_stage(vsb, ref_lib, clip, _corrections(n, 6, 6.0), kw)
    assert max(filled) == 3


def test_stage_small_canvas_fills_the_surround(vsb):
    """canvas 1.25x / 1.3x: the surround's bounding rectangle (the whole canvas) is fillable: warp + cut-out + resize to the
    canvas + blend over everything, current frame included; short temporal buffer, narrow edge ramp; padded rows"""
    ref_lib = _ref_lib()
    w, h, n = 322, 182, 12
    clip = _bright_clip(w, h, n, 4)
    clip[:, 40:60, 40:70] = 0
    kw = dict(canvasScaleFactor=1.25, adaptiveCanvasSize=False, temporalBufferSize=4, edgeBlendRadius=7, canvasBlendWeight=0.45)
    filled, _ = _run_stage(vsb, ref_lib, clip, _corrections(n, 2, 25.0), kw, stride_pad=10)
    assert max(filled) == 1
    kw = dict(canvasScaleFactor=1.3, adaptiveCanvasSize=False)
    filled, _ = _run_stage(vsb, ref_lib, _bright_clip(320, 180, 10, 5), _corrections(10, 3, 45.0, rot=0.05), kw)
    assert max(filled) == 1


@pytest.mark.parametrize("mag,want", [(70.0, None), (130.0, 2.0), (10.0, 1.5)])
def test_stage_adaptive_canvas_size(vsb, mag, want):
    """calculateOptimalCanvasSize (:2280-2314): the largest of the last 30 transforms sizes the canvas, once"""
    ref_lib = _ref_lib()
    rng = np.random.default_rng(7)
    recent = np.zeros((40, 3), np.float32)
    recent[:, 0] = rng.uniform(-1, 1, 40) * mag * 0.7
    recent[:, 1] = rng.uniform(-1, 1, 40) * mag * 0.7
    recent[3, :2] = 500.0                            # older than the last 30: not looked at
    clip = _bright_clip(320, 180, 6, 6)
    clip[:, 80:100, 150:180] = 0
    _, scale = _run_stage(vsb, ref_lib, clip, _corrections(6, 4, 60.0), dict(minCanvasScale=1.0), recent=recent)
    if want is not None:
        assert scale == want
    else:
        assert 1.5 < scale < 2.0


def test_stage_many_regions_and_black_frames(vsb):
    """more regions than one launch takes (40 dark boxes), overlapping bounding rectangles (an L-shaped object around a box),
    a completely black frame and a frame with one bright pixel"""
    ref_lib = _ref_lib()
    w, h, n = 480, 270, 9
    clip = _bright_clip(w, h, n, 8)
    for j in range(5):
        for i in range(8):
            clip[:, 20 + j * 48: 32 + j * 48, 20 + i * 56: 33 + i * 56] = 0
    clip[:, 200:260, 300:310] = 0
    clip[:, 250:260, 300:380] = 0                    # L: its bounding rectangle contains the box below
    clip[:, 215:235, 330:360] = 0
    clip[5] = 0
    clip[6] = 0
    clip[6, 100, 100] = 255
    filled, _ = _run_stage(vsb, ref_lib, clip, _corrections(n, 5, 12.0), dict(canvasScaleFactor=1.0, adaptiveCanvasSize=False))
    assert max(filled) > 24


def test_stabilizer_with_virtual_canvas_vs_reference(vsb):
    """enableVirtualCanvas inside the stabilizer: frame-sized output even with a border, the last frame passed through.  The
    stage moves the frame by the INTEGER part of a correction that the device and the reference compute to within 1e-3 px
    (DESIGN.md 2): frames are bit-exact unless a correction sits on an integer boundary."""
    ref_lib = _ref_lib()
    w, h, n = 640, 360, 30
    clip = _bright_clip(w, h, n, 9)
    clip[:, 100:130, 200:260] = 0
    for kw in (dict(smoothingRadius=5), dict(smoothingRadius=6, borderSize=16, borderType="reflect", canvasScaleFactor=1.25, adaptiveCanvasSize=False)):
        P = dict(enableVirtualCanvas=True, **kw)
        ref = ref_lib.RefStabilizer(P)
        st = vsb.Stabilizer(vsb.Parameters(**P))
        exact = total = 0
        for k in range(n + 40):
            if k < n:
                want, got = ref.stabilize(clip[k]), st.stabilize(clip[k])
            else:
                want, got = ref.flush(), st.flush()
                if want is None:
                    assert got is None
                    break
            assert (want is None) == (got is None)
            if want is None:
                continue
            assert want.shape == got.shape == (h, w, 3)
            total += 1
            if np.array_equal(want, got):
                exact += 1
            else:
                t, _, _ = ref.vc_last()
                rec = st.output_record(total - 1)
                mine = np.array([rec.T[2], rec.T[5]], np.float32)
                assert np.abs(mine - t[:2]).max() < 2e-3, f"frame {total - 1}: corrections {mine} vs {t[:2]}"
        assert total == n and exact >= total - 3, f"{exact} of {total} frames bit-exact"


def test_virtual_canvas_is_skipped_with_crop_n_zoom_and_refused_where_it_cannot_run(vsb):
    clip = _bright_clip(320, 180, 12, 10)
    a = vsb.Stabilizer(vsb.Parameters(smoothingRadius=5, cropNZoom=True, borderSize=10))
    b = vsb.Stabilizer(vsb.Parameters(smoothingRadius=5, cropNZoom=True, borderSize=10, enableVirtualCanvas=True))
    for f in clip:
        x, y = a.stabilize(f), b.stabilize(f)
        assert (x is None) == (y is None) and (x is None or np.array_equal(x, y))      # :1110-1127 return before the stage
    with pytest.raises(vsb.VsError):
        vsb.StabilizerBatch(vsb.Parameters(enableVirtualCanvas=True), 2)
    with pytest.raises(vsb.VsError):
        vsb.Stabilizer(vsb.Parameters(enableVirtualCanvas=True, canvasScaleFactor=0.8, adaptiveCanvasSize=False))
    with pytest.raises(vsb.VsError):
        vsb.VirtualCanvas(vsb.Parameters(enableVirtualCanvas=True, temporalBufferSize=-1))

"""Offline clip mode (temporal chunking + all-gather stitch, BASELINE config 5)."""
import os
import sys

import numpy as np

import synthclip
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_chunk_bounds_cover_the_clip():
    import __graft_entry__
    __graft_entry__.build()
    from video_stab_b200 import offline
    for n in (1, 2, 7, 48, 300, 18000):
        for world in (1, 2, 3, 4, 8):
            cur = 0
            for r in range(world):
                f, c = offline.chunk_bounds(n, world, r)
                assert f == min(cur, n) and c >= 0
                assert f % 2 == 0 or c == 0
                cur = f + c
            assert cur == n
    assert [offline.halo(f) for f in (0, 1, 2, 3, 4, 5, 6, 100, 101)] == [0, 1, 2, 1, 2, 1, 2, 2, 1]


def _stitch_worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from video_stab_b200 import offline
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    full = np.arange((n_total - 1) * 3, dtype=np.float32).reshape(-1, 3) * 0.25
    f, c = offline.chunk_bounds(n_total, world, rank)
    lo, hi = max(f, 1) - 1, f + c - 1
    local = full[lo:hi] if c > 0 else full[:0]
    got = offline.stitch_transforms(local, f, c, n_total)
    q.put((rank, bool(np.array_equal(got, full))))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [37, 300])
def test_stitch_all_gather_gloo_world2(n_total):
    """The N>1 exchange step on CPU: two ranks, gloo backend, ragged chunks."""
    import torch.multiprocessing as mp
    import __graft_entry__
    __graft_entry__.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_total % 7
    procs = [ctx.Process(target=_stitch_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(smoothingRadius=8), dict(smoothingRadius=6, smoothingMethod="gaussian"),
                                dict(smoothingRadius=6, smoothingMethod="kalman"),
                                dict(smoothingRadius=5, borderType="reflect", borderSize=16),
                                dict(smoothingRadius=5, cropNZoom=True, borderSize=20)])
def test_chunked_clip_equals_streamed(kw):
    """Chunked (1, 2, 3 and 5 chunks) == streamed stabilize()+flush(), bit for bit: same transforms, same frames."""
    import torch
    import video_stab_b200 as vsb
    w, h, n = 640, 360, 41
    clip = synthclip.make_clip(w, h, n, 321)
    params = vsb.Parameters(**kw)
    st = vsb.Stabilizer(params)
    outs = [o for o in (st.stabilize(f) for f in clip) if o is not None]
    while True:
        o = st.flush()
        if o is None:
            break
        outs.append(o)
    ref_tr = np.array([list(st.frame_record(i).transform) for i in range(n - 1)], np.float32)
    d = torch.from_numpy(clip).cuda()
    for chunks in (1, 2, 3, 5):
        got, tr = vsb.offline.stabilize_clip(d, params, n_chunks=chunks)
        assert np.array_equal(tr.view(np.uint32), ref_tr.view(np.uint32)), f"{chunks} chunks: transforms differ"
        g = got.cpu().numpy()
        for i in range(n):
            oh, ow = outs[i].shape[:2]
            assert np.array_equal(g[i, :oh, :ow], outs[i]), f"{chunks} chunks: frame {i} differs"

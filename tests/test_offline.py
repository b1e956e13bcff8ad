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


def test_lockstep_chunking_gives_equal_even_chunks():
    """offline.lockstep_chunking: equal chunks of an even number of frames (>= 40), at most 64 per rank, for the clip lengths and
    world sizes of BASELINE config 5; every chunk but the clip's first starts at an even frame >= 4 with a two-frame halo."""
    import __graft_entry__
    __graft_entry__.build()
    from video_stab_b200 import offline
    for n in (18000, 4096, 9000, 600):
        for world in (1, 2, 4, 8):
            cpr = offline.lockstep_chunking(n, world)
            if cpr is None:
                continue
            chunks = world * cpr
            per = n // chunks
            assert 2 <= cpr <= 64 and per * chunks == n and per % 2 == 0 and per >= 40
            per_s, idx = offline.stitch_index(n, chunks)
            assert per_s == per and len(idx) == n - 1
            for c in range(chunks):
                first, count = offline.chunk_bounds(n, chunks, c)
                assert count == per and first == c * per
                if c:
                    assert first % 2 == 0 and first >= 4 and offline.halo(first) == 2
    assert offline.lockstep_chunking(18000, 1) == 60 and offline.lockstep_chunking(18000, 8) == 45
    assert offline.lockstep_chunking(37, 2) is None


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
    # the tensor variant (what the device-resident path uses; here on CPU tensors over gloo)
    import torch
    per, _ = offline.stitch_index(n_total, world)
    buf = torch.full((per, 3), -1.0)
    buf[: len(local)] = torch.from_numpy(np.ascontiguousarray(local))
    got2 = offline.stitch_transforms_tensor(buf, n_total).numpy()
    q.put((rank, bool(np.array_equal(got, full)) and bool(np.array_equal(got2, full))))
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


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n,world,method", [(640, 360, 480, 1, "box"), (1280, 720, 960, 2, "box"),
                                                 (640, 360, 480, 1, "gaussian"), (640, 360, 480, 1, "kalman")])
def test_lockstep_chunk_analysis_equals_streamed(w, h, n, world, method):
    """The lock-step analysis (vs_batch_clip_analyze_device: all of a rank's chunks advance together, one launch per stage)
    and the prepared render (vs_clip_set_transforms_device once, vs_clip_render_prepared_device per chunk) against the
    streamed stabilize() + flush() of the whole clip: transforms bit for bit, every output frame equal.  `world` ranks are
    played one after the other on this GPU (rank 0 owns the clip's first chunk, which goes frame by frame)."""
    import torch
    import video_stab_b200 as vsb
    from video_stab_b200 import offline
    dev = torch.device("cuda", 0)
    gen = synthclip.DeviceClip(w, h, n, 777, dev)
    clip = gen.frames(0, n)
    params = vsb.Parameters(smoothingRadius=9, smoothingMethod=method)       # (kalman: the recursion continues from chunk to chunk)
    fb = h * w * 3
    st = vsb.Stabilizer(params)
    ring = torch.empty((n + 1, h, w, 3), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    k = st.push_many_device(clip.data_ptr(), fb, n, w, h, w * 3, ring.data_ptr(), w * 3, fb, borrow=False)
    while st.flush_device(ring[k].data_ptr(), w * 3, fb) is not None:
        k += 1
    st.sync()
    assert k == n
    ref_tr = np.array([list(st.frame_record(i).transform) for i in range(n - 1)], np.float32)
    del st

    import ctypes as C
    cpr = offline.lockstep_chunking(n, world, max_lanes=12, min_frames=20)
    assert cpr is not None
    chunks = world * cpr
    per, idx = offline.stitch_index(n, chunks)

    def chunk_tensor(c):
        first, count = offline.chunk_bounds(n, chunks, c)
        return clip[first - offline.halo(first): first + count]

    if world == 1:
        mine = list(range(chunks))
        frames = {c: chunk_tensor(c) for c in mine}
        st = vsb.Stabilizer(params)
        sb = vsb.StabilizerBatch(params, chunks - 1)                  # every chunk but the clip's first advances in lock-step
        out = torch.zeros((per, h, w, 3), dtype=torch.uint8, device=dev)
        full = offline.stabilize_rank_chunks(st, frames, n, chunks, mine, out, batch=sb)
        st.sync(); sb.sync(); torch.cuda.synchronize()
        assert np.array_equal(full.cpu().numpy().view(np.uint32), ref_tr.view(np.uint32)), "lock-step transforms differ from streamed"
        last_first, last_count = offline.chunk_bounds(n, chunks, chunks - 1)
        assert torch.equal(out[:last_count], ring[last_first:last_first + last_count]), "last chunk: frames differ"
        # every chunk rendered against the installed transform list equals the streamed output
        ow, oh = C.c_int(), C.c_int()
        for c in mine:
            first, count = offline.chunk_bounds(n, chunks, c)
            offline.check(offline.lib.vs_clip_render_prepared_device(st._h, clip[first].data_ptr(), w, h, first, count,
                                                                     out.data_ptr(), C.byref(ow), C.byref(oh)))
            st.sync()
            assert torch.equal(out[:count], ring[first:first + count]), f"chunk {c}: frames differ"
    else:
        # `world` ranks played in turn: each rank's lock-step rows must equal the streamed transforms of its frames
        for rank in range(world):
            mine = [c for c in range(rank * cpr, (rank + 1) * cpr) if offline.chunk_bounds(n, chunks, c)[0] >= 4]
            frames = {c: chunk_tensor(c) for c in mine}
            sb = vsb.StabilizerBatch(params, len(mine))
            local = torch.zeros((len(mine) * per, 3), dtype=torch.float32, device=dev)
            torch.cuda.synchronize()
            sb.clip_analyze_device([frames[c].data_ptr() for c in mine], w, h, per, [local[k * per].data_ptr() for k in range(len(mine))])
            sb.sync()
            loc = local.cpu().numpy()
            for k, c in enumerate(mine):
                first, count = offline.chunk_bounds(n, chunks, c)
                assert np.array_equal(loc[k * per: k * per + count].view(np.uint32), ref_tr[first - 1:first + count - 1].view(np.uint32)), \
                    f"rank {rank} chunk {c}: lock-step transforms differ"


def _synthetic_long_clip(torch, w, h, n, seed, dev):
    """n frames generated on the device from the seed (synthclip.DeviceClip); returns frames(a, b)."""
    return synthclip.DeviceClip(w, h, n, seed, dev).frames


def _frame_checksums(torch, frames):
    """per-frame 64-bit checksum on the device (position-weighted sum of the 32-bit words)"""
    n = frames.shape[0]
    words = frames.reshape(n, -1).view(torch.int32).to(torch.int64)
    wts = (torch.arange(words.shape[1], device=frames.device, dtype=torch.int64) % 65521) + 1
    return (words * wts).sum(1)


@pytest.mark.gpu
def test_config5_long_clip_chunked_equals_streamed(monkeypatch):
    """BASELINE config 5 at its full length: an 18 000-frame 1080p clip (10 min @ 30 fps), generated on the device.
    8 temporal chunks through the device-resident clip calls (analyse -> stitch -> rebuild path -> smooth -> warp) must
    equal the streamed stabilize() + flush() of the whole clip: transforms bit for bit, every output frame by checksum.
    The trajectory arrays start at 4096 entries (VS_TRAJ_CAP), so grow_trajectory() runs three times on the way."""
    import torch
    import video_stab_b200 as vsb
    from video_stab_b200 import offline
    monkeypatch.setenv("VS_TRAJ_CAP", "4096")
    w, h, n, chunks = 1920, 1080, 18000, 8
    dev = torch.device("cuda", 0)
    gen = _synthetic_long_clip(torch, w, h, n, 5000, dev)
    params = vsb.Parameters(smoothingRadius=15)
    fb = h * w * 3

    # ---- streamed: push_device frame by frame, outputs into a 512-frame ring, checksummed block by block
    st = vsb.Stabilizer(params)
    blk = 500
    ring = torch.empty((blk + 40, h, w, 3), dtype=torch.uint8, device=dev)
    sums = []
    produced = 0
    for a in range(0, n, blk):
        fr = gen(a, min(a + blk, n))
        torch.cuda.synchronize()
        k = st.push_many_device(fr.data_ptr(), fb, fr.shape[0], w, h, w * 3, ring.data_ptr(), w * 3, fb, borrow=False)
        st.sync()
        sums.append(_frame_checksums(torch, ring[:k]).cpu())
        produced += k
    while True:
        got = st.flush_device(ring.data_ptr(), w * 3, fb)
        if got is None:
            break
        st.sync()
        sums.append(_frame_checksums(torch, ring[:1]).cpu())
        produced += 1
    assert produced == n
    streamed = torch.cat(sums)
    nf, no = st.counts()
    assert nf == n - 1 and no == n
    ref_tr = np.array([list(st.frame_record(i).transform) for i in range(0, n - 1, 97)], np.float32)
    last = st.output_record(n - 2)
    del st

    # ---- chunked: 8 chunks on this one GPU, one after the other, exactly as 8 ranks would run them
    st = vsb.Stabilizer(params)
    per, idx = offline.stitch_index(n, chunks)
    gathered = torch.zeros((chunks * per, 3), dtype=torch.float32, device=dev)
    for r in range(chunks):
        first, count = offline.chunk_bounds(n, chunks, r)
        hl = offline.halo(first)
        fr = gen(first - hl, first + count)
        torch.cuda.synchronize()
        offline.analyze_chunk_device(st, fr.data_ptr(), w, h, first, count, gathered[r * per].data_ptr())
        st.sync()
        del fr
    full = gathered.index_select(0, torch.from_numpy(idx).to(dev)).contiguous()
    tr = full.cpu().numpy()
    assert np.array_equal(tr[::97].view(np.uint32), ref_tr.view(np.uint32)), "chunked transforms differ from streamed"
    chunked = []
    for r in range(chunks):
        first, count = offline.chunk_bounds(n, chunks, r)
        fr = gen(first, first + count)
        out = torch.empty((count, h, w, 3), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        offline.render_chunk_device(st, full.data_ptr(), n, fr.data_ptr(), w, h, first, count, out.data_ptr())
        st.sync()
        for a in range(0, count, 250):
            chunked.append(_frame_checksums(torch, out[a:a + 250]).cpu())
        del fr, out
    chunked = torch.cat(chunked)
    bad = (chunked != streamed).nonzero().flatten().tolist()
    assert not bad, f"{len(bad)} of {n} output frames differ, first at {bad[:5]}"
    assert last.index == n - 2

"""Offline clip mode (BASELINE config 5): a long clip is split into contiguous temporal chunks, one per
GPU.  Each rank analyses its own frames (plus a <= 2 frame pixel halo), the per-frame transforms are
exchanged with ONE tiny all-gather (12 bytes per frame), every rank rebuilds the whole trajectory in the
reference's sequential float32 order, then smooths and warps its own frames in batched launches.  The
result is bit-identical to pushing the whole clip through `Stabilizer.stabilize()` + `flush()`.

`torch.distributed` is only the plumbing for the all-gather (NCCL over NVLink on GPUs, gloo in the CPU
tests); all arithmetic is in the CUDA library."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._capi import check, lib
from .stabilizer import Parameters, Stabilizer


def chunk_bounds(n_frames: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous chunk [first, first+count) of rank `rank`; chunk starts are kept even so every chunk
    needs the same 2-frame halo (corners come from the even frame before it)."""
    per = -(-n_frames // world)
    per += per & 1
    first = min(rank * per, n_frames)
    return first, max(0, min(per, n_frames - first))


def halo(first: int) -> int:
    return lib.vs_clip_halo(first)


def analyze_chunk(st: Stabilizer, d_frames_ptr: int, w: int, h: int, first: int, count: int) -> np.ndarray:
    """Transforms of the generateTransform calls n = max(first,1)..first+count-1, shape (k,3) float32.
    `d_frames_ptr`: device address of frame `first - halo(first)` (tight rows, frames contiguous)."""
    out = np.zeros((max(count, 1), 3), np.float32)
    n = C.c_int()
    check(lib.vs_clip_analyze(st._h, d_frames_ptr, w, h, first, count, out.ctypes.data, C.byref(n)))
    return out[: n.value].copy()


def render_chunk(st: Stabilizer, all_transforms: np.ndarray, n_total: int, d_frames_ptr: int, w: int, h: int,
                 first: int, count: int, d_out_ptr: int) -> tuple[int, int]:
    tr = np.ascontiguousarray(all_transforms, np.float32).reshape(-1, 3)
    assert len(tr) == n_total - 1
    ow, oh = C.c_int(), C.c_int()
    check(lib.vs_clip_render(st._h, tr.ctypes.data, n_total, d_frames_ptr, w, h, first, count, d_out_ptr,
                             C.byref(ow), C.byref(oh)))
    return ow.value, oh.value


def stitch_transforms(local: np.ndarray, first: int, count: int, n_total: int, group=None) -> np.ndarray:
    """All-gather of the per-chunk transforms -> the (n_total-1, 3) transform list of the whole clip.
    Every rank contributes transforms_[max(first,1)-1 .. first+count-2]; payload 12 bytes per frame."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        assert len(local) == n_total - 1
        return np.ascontiguousarray(local, np.float32)
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    per = max(chunk_bounds(n_total, world, r)[1] for r in range(world))
    buf = torch.zeros((per, 3), dtype=torch.float32, device=dev)
    if len(local):
        buf[: len(local)] = torch.from_numpy(np.ascontiguousarray(local, np.float32)).to(dev)
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf, group=group)
    parts = []
    for r in range(world):
        f, c = chunk_bounds(n_total, world, r)
        k = f + c - max(f, 1) if c > 0 else 0
        parts.append(gathered[r][:k].cpu().numpy())
    out = np.concatenate(parts, axis=0) if parts else np.zeros((0, 3), np.float32)
    assert len(out) == n_total - 1, (len(out), n_total)
    return out


# ------------------------------------------------------------------------------------------ device-resident path
# The transforms never leave the device: vs_clip_analyze_device writes them to a device buffer, the all-gather (NCCL
# over NVLink) exchanges that buffer in place, vs_clip_render_device consumes the stitched list.  Nothing blocks the host.
def stitch_index(n_total: int, world: int):
    """Row indices into the (world * per, 3) all-gather result that give the clip's n_total-1 transforms in order
    (rank r contributed its k_r = first+count-max(first,1) transforms at rows r*per ..)."""
    per = max(chunk_bounds(n_total, world, r)[1] for r in range(world))
    idx = []
    for r in range(world):
        f, c = chunk_bounds(n_total, world, r)
        k = f + c - max(f, 1) if c > 0 else 0
        idx.extend(range(r * per, r * per + k))
    assert len(idx) == n_total - 1, (len(idx), n_total)
    return per, np.asarray(idx, np.int64)


def stitch_transforms_tensor(local, n_total: int, group=None):
    """`local`: (per, 3) float32 tensor on the rank's device (rows beyond the rank's own transforms are ignored).
    Returns the (n_total-1, 3) transform list of the whole clip on the same device.  One all-gather of 12 bytes per
    frame (NCCL on GPUs, gloo in the CPU tests); no host round trip."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    per, idx = stitch_index(n_total, world)
    assert tuple(local.shape) == (per, 3) and local.dtype == torch.float32
    if world == 1:
        gathered = local
    elif dist.get_backend(group) == "nccl":
        gathered = torch.empty((world * per, 3), dtype=torch.float32, device=local.device)
        dist.all_gather_into_tensor(gathered, local.contiguous(), group=group)
    else:
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local.contiguous(), group=group)
        gathered = torch.cat(parts, 0)
    return gathered.index_select(0, torch.from_numpy(idx).to(local.device))


def analyze_chunk_device(st: Stabilizer, d_frames_ptr: int, w: int, h: int, first: int, count: int, d_out_ptr: int) -> int:
    """Asynchronous vs_clip_analyze: transforms land in device memory at d_out_ptr, ordered on st.stream."""
    n = C.c_int()
    check(lib.vs_clip_analyze_device(st._h, d_frames_ptr, w, h, first, count, d_out_ptr, C.byref(n)))
    return n.value


def render_chunk_device(st: Stabilizer, d_all_transforms_ptr: int, n_total: int, d_frames_ptr: int, w: int, h: int,
                        first: int, count: int, d_out_ptr: int) -> tuple[int, int]:
    ow, oh = C.c_int(), C.c_int()
    check(lib.vs_clip_render_device(st._h, d_all_transforms_ptr, n_total, d_frames_ptr, w, h, first, count, d_out_ptr,
                                    C.byref(ow), C.byref(oh)))
    return ow.value, oh.value


def stabilize_chunk_distributed(st: Stabilizer, frames, halo_frames: int, n_total: int, first: int, count: int, out, group=None):
    """One rank's share of a long clip, end to end on the device: analyse -> all-gather -> rebuild path -> smooth -> warp.
    `frames`: (halo_frames + count, H, W, 3) uint8 CUDA tensor holding frames [first - halo_frames, first + count);
    `out`: (count, H', W', 3) uint8 CUDA tensor.  Asynchronous; returns the stitched (n_total-1, 3) transform tensor."""
    import torch
    import torch.distributed as dist
    assert halo_frames == halo(first)
    h, w = int(frames.shape[1]), int(frames.shape[2])
    fb = h * w * 3
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    per, _ = stitch_index(n_total, world)
    local = torch.zeros((per, 3), dtype=torch.float32, device=frames.device)
    ext = torch.cuda.ExternalStream(st.stream, device=frames.device)
    cur = torch.cuda.current_stream(frames.device)
    if count > 0:
        analyze_chunk_device(st, frames.data_ptr(), w, h, first, count, local.data_ptr())
    ev = torch.cuda.Event()
    ev.record(ext)
    cur.wait_event(ev)
    full = stitch_transforms_tensor(local, n_total, group)
    ev2 = torch.cuda.Event()
    ev2.record(cur)
    st.wait_event(ev2.cuda_event)
    if count > 0:
        render_chunk_device(st, full.data_ptr(), n_total, frames.data_ptr() + halo_frames * fb, w, h, first, count, out.data_ptr())
    return full


def lockstep_chunking(n_total: int, world: int, max_lanes: int = 64, min_frames: int = 40):
    """Number of temporal chunks per rank for the lock-step analysis: as many (<= max_lanes) as divide the clip into equal
    chunks of an even number (>= min_frames) of frames - equal chunks advance in lock-step, even starts keep the corner
    re-detection (every second frame) aligned across chunks.  None if the clip length does not divide that way."""
    for cpr in range(max_lanes, 1, -1):
        n_chunks = world * cpr
        if n_total % n_chunks == 0:
            per = n_total // n_chunks
            if per % 2 == 0 and per >= min_frames:
                return cpr
    return None


def stabilize_rank_chunks(st: Stabilizer, chunk_frames: dict, n_total: int, n_chunks: int, my_chunks, out_ring, group=None,
                          batch=None):
    """One rank's share of a long clip cut into `n_chunks` temporal chunks (n_chunks a multiple of the world size; rank r
    owns the contiguous block `my_chunks`).  chunk_frames[c]: (halo + count, H, W, 3) uint8 CUDA tensor of frames
    [first_c - halo_c, first_c + count_c).  Every chunk is analysed, ONE all-gather stitches the transforms of the whole
    clip (device to device), then every chunk is smoothed and warped into `out_ring` ((max count, H', W', 3); reused chunk
    after chunk when a rank owns several).  With `batch` (a StabilizerBatch) the rank's chunks that start at an even frame
    >= 4 and have the common length are analysed in LOCK-STEP, batch.n at a time (one launch per stage for all of them);
    the rest (the clip's first chunk, a shorter last one) go through `st` frame by frame, concurrently.  Asynchronous on
    the handles' streams and torch's current stream; returns the stitched (n_total-1, 3) transform tensor."""
    import torch
    import torch.distributed as dist
    my_chunks = list(my_chunks)
    per, idx = stitch_index(n_total, n_chunks)
    some = chunk_frames[my_chunks[0]]
    dev, h, w = some.device, int(some.shape[1]), int(some.shape[2])
    fb = h * w * 3
    local = torch.zeros((len(my_chunks) * per, 3), dtype=torch.float32, device=dev)
    ext = torch.cuda.ExternalStream(st.stream, device=dev)
    cur = torch.cuda.current_stream(dev)
    ev0 = torch.cuda.Event()
    ev0.record(cur)
    st.wait_event(ev0.cuda_event)                      # `local` was zeroed on torch's stream
    lock = []
    if batch is not None:
        batch.wait_event(ev0.cuda_event)
        for k, c in enumerate(my_chunks):
            first, count = chunk_bounds(n_total, n_chunks, c)
            if first >= 4 and first % 2 == 0 and count == per and halo(first) == 2:
                lock.append((k, c))
        lock = lock[: (len(lock) // batch.n) * batch.n] if len(lock) >= batch.n else []
    in_lock = {c for _, c in lock}
    for r0 in range(0, len(lock), batch.n if batch is not None else 1):
        grp = lock[r0:r0 + batch.n]
        batch.clip_analyze_device([chunk_frames[c].data_ptr() for _, c in grp], w, h, per,
                                  [local[k * per].data_ptr() for k, _ in grp])
    for k, c in enumerate(my_chunks):
        first, count = chunk_bounds(n_total, n_chunks, c)
        if count > 0 and c not in in_lock:
            analyze_chunk_device(st, chunk_frames[c].data_ptr(), w, h, first, count, local[k * per].data_ptr())
    ev = torch.cuda.Event()
    ev.record(ext)
    cur.wait_event(ev)
    if lock:
        evb = torch.cuda.Event()
        evb.record(torch.cuda.ExternalStream(batch.stream, device=dev))
        cur.wait_event(evb)
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if world > 1:
        gathered = torch.empty((n_chunks * per, 3), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(gathered, local, group=group)
    else:
        gathered = local
    full = gathered.index_select(0, torch.from_numpy(idx).to(dev))
    ev2 = torch.cuda.Event()
    ev2.record(cur)
    st.wait_event(ev2.cuda_event)
    check(lib.vs_clip_set_transforms_device(st._h, full.data_ptr(), n_total, w, h))      # trajectory rebuilt once per clip
    ow, oh = C.c_int(), C.c_int()
    for c in my_chunks:
        first, count = chunk_bounds(n_total, n_chunks, c)
        if count > 0:
            hl = halo(first)
            check(lib.vs_clip_render_prepared_device(st._h, chunk_frames[c].data_ptr() + hl * fb, w, h, first, count,
                                                     out_ring.data_ptr(), C.byref(ow), C.byref(oh)))
    return full


def stabilize_clip(frames, params: Parameters | None = None, out=None, n_chunks: int = 1, device: int = 0):
    """Single-process driver: `frames` is a (N,H,W,3) uint8 CUDA tensor.  With n_chunks > 1 the clip is
    processed chunk by chunk exactly as n_chunks ranks would (used to test the chunked path on one GPU)."""
    import torch
    params = params or Parameters()
    n, h, w, _ = frames.shape
    b = params.borderSize if (params.borderSize > 0 and not params.cropNZoom) else 0
    if out is None:
        out = torch.zeros((n, h + 2 * b, w + 2 * b, 3), dtype=torch.uint8, device=frames.device)
    st = Stabilizer(params, device=device)
    fb = h * w * 3
    parts = []
    for r in range(n_chunks):
        first, count = chunk_bounds(n, n_chunks, r)
        if count == 0:
            continue
        hl = halo(first)
        parts.append(analyze_chunk(st, frames.data_ptr() + (first - hl) * fb, w, h, first, count))
    tr = np.concatenate(parts, axis=0) if parts else np.zeros((0, 3), np.float32)
    for r in range(n_chunks):
        first, count = chunk_bounds(n, n_chunks, r)
        if count == 0:
            continue
        render_chunk(st, tr, n, frames.data_ptr() + first * fb, w, h, first, count, out[first].data_ptr())
    torch.cuda.synchronize()
    return out, tr

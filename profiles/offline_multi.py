"""BASELINE config 5 on N GPUs: one temporal chunk per rank, ONE all-gather of the per-frame transforms (NCCL), every
rank renders its own frames.  Checks the stitched result against a 1-GPU run of the whole clip on rank 0 (bit-exact
transforms and frames) and prints device-timed throughput (max over ranks).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 \
        profiles/offline_multi.py [n_frames]"""
import json
import os
import sys
import time

import numpy as np

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.build()
import video_stab_b200 as vsb  # noqa: E402
import synthclip
from video_stab_b200 import offline  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
W, H = 1920, 1080
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
fb = W * H * 3
base = torch.from_numpy(synthclip.make_clip(W, H, 64, 5000)).to(dev)
pp = list(range(64)) + list(range(62, 0, -1))


def frames(lo, hi):            # the long clip is generated on the fly from the seed clip (a 10-minute raw clip is 112 GB)
    return base[torch.tensor([pp[k % 126] for k in range(lo, hi)], device=dev)]


params = vsb.Parameters(smoothingRadius=15)
first, count = offline.chunk_bounds(n, world, rank)
hl = offline.halo(first)
mine = frames(first - hl, first + count)
out = torch.empty((count, H, W, 3), dtype=torch.uint8, device=dev)
torch.cuda.synchronize()        # the handle's streams do not order against torch's stream: the frames must be complete
st = vsb.Stabilizer(params, device=local)


def run():
    local_tr = offline.analyze_chunk(st, mine.data_ptr(), W, H, first, count)
    tr = offline.stitch_transforms(local_tr, first, count, n)
    offline.render_chunk(st, tr, n, mine.data_ptr() + hl * fb, W, H, first, count, out.data_ptr())
    return tr


run()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
tr = run()
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
crc = torch.tensor([int(out.view(-1).to(torch.int64).sum().item())], dtype=torch.int64, device=dev)
sums = [torch.zeros_like(crc) for _ in range(world)]
if world > 1:
    dist.all_gather(sums, crc)
else:
    sums = [crc]
ok = None
if rank == 0:
    # reference: the whole clip on this GPU in one chunk
    st1 = vsb.Stabilizer(params, device=local)
    full = frames(0, n)
    torch.cuda.synchronize()
    tr1 = offline.analyze_chunk(st1, full.data_ptr(), W, H, 0, n)
    ok_tr = bool(np.array_equal(tr1.view(np.uint32), tr.view(np.uint32)))
    if not ok_tr:
        bad = np.nonzero((tr1.view(np.uint32) != tr.view(np.uint32)).any(axis=1))[0]
        print("transform rows that differ:", bad[:16], "of", len(bad), "chunk size", offline.chunk_bounds(n, world, 0)[1], file=sys.stderr)
        for i in bad[:4]:
            print(i, tr1[i], tr[i], file=sys.stderr)
    ref_out = torch.empty_like(full)
    offline.render_chunk(st1, tr1, n, full.data_ptr(), W, H, 0, n, ref_out.data_ptr())
    ok_frames = True
    for r in range(world):
        f, c = offline.chunk_bounds(n, world, r)
        ok_frames = ok_frames and int(ref_out[f:f + c].view(-1).to(torch.int64).sum().item()) == int(sums[r].item())
    ok_mine = bool(torch.equal(ref_out[first:first + count], out))
    print(json.dumps({"config": "offline clip, temporal chunks + all-gather stitch", "n_gpus": world, "frames": n,
                      "seconds": float(dt), "frames_per_s": n / float(dt), "transforms_bit_exact": ok_tr,
                      "rank0_frames_bit_exact": ok_mine, "all_rank_frame_sums_match": ok_frames}))
if world > 1:
    dist.destroy_process_group()

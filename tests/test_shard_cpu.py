"""Host-side sharding logic on the CPU: the planners, and a world_size-2 gloo run that
checks a frame-range sharded causal computation against the single-process result."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from compressd_domain_saliencyprediction_b200 import shard


def test_halo_matches_reference_windows():
    assert shard.halo_frames(50.0) == 15 + 105 - 2       # main.m:119,129 at 50 fps
    assert shard.halo_frames(24.0) == 7 + 49 - 2
    assert shard.halo_frames(60.0) == 18 + 126 - 2


def test_plan_videos_balances_and_covers():
    lengths = [600, 500, 500, 300, 240, 150, 601, 500]
    plan = shard.plan_videos(lengths, 4)
    assert sorted(i for p in plan for i in p) == list(range(len(lengths)))
    loads = [sum(lengths[i] for i in p) for p in plan]
    assert max(loads) - min(loads) <= max(lengths)
    assert shard.plan_videos([5], 3) == [[0], [], []]


def test_plan_frame_ranges_partition_with_halo():
    for nframes, world, halo in [(500, 8, 118), (10, 4, 22), (3, 4, 5), (1000, 2, 54)]:
        plan = shard.plan_frame_ranges(nframes, world, halo)
        assert sum(p["count"] for p in plan) == nframes
        nxt = 0
        for p in plan:
            assert p["first"] == nxt and p["seek"] == max(0, p["first"] - halo) and p["context"] == p["first"] - p["seek"]
            nxt += p["count"]


def _causal(frames, halo):
    """Stand-in for the temporal dependence of the real path: out[k] depends on frames k-halo..k."""
    out = np.zeros(len(frames))
    for k in range(len(frames)):
        lo = max(0, k - halo)
        out[k] = np.sum(frames[lo:k + 1] * np.exp(-np.arange(k - lo, -1, -1) / 7.0))
    return out


def _worker(rank, world, port, nframes, halo, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = np.random.default_rng(11).standard_normal(nframes)
    p = shard.plan_frame_ranges(nframes, world, halo)[rank]
    # a rank sees only seek .. first+count-1, exactly what cds_seek + cds_clip_context + cds_clip_process consume
    window = frames[p["seek"]:p["first"] + p["count"]]
    local = _causal(window, halo)[p["context"]:]
    full = shard.gather_to_rank0(local.reshape(-1, 1), dist)
    if rank == 0:
        q.put(full[:, 0])
    dist.barrier()
    dist.destroy_process_group()


def test_frame_range_sharding_world2_gloo():
    nframes, halo, world = 97, 22, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nframes, halo, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _causal(np.random.default_rng(11).standard_normal(nframes), halo)
    assert np.array_equal(got, want)         # same arithmetic on the same window -> bytewise identical


def test_halo_of_the_fast_version():
    """salmat.cpp:18-19: cvRound(0.3 fps) + cvRound(0.7 fps) - 2 pictures of context (cvRound rounds halves to even)."""
    from compressd_domain_saliencyprediction_b200 import shard
    assert shard.halo_frames(50.0, fast=True) == 15 + 35 - 2
    assert shard.halo_frames(25.0, fast=True) == 8 + 18 - 2          # 7.5 -> 8, 17.5 -> 18
    assert shard.halo_frames(30.0, fast=True) == 9 + 21 - 2
    assert shard.halo_frames(50.0) == 15 + 105 - 2

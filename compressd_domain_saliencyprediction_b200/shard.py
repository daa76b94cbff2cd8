"""Host-side partitioning of the hot path over the GPUs of one box.

The path shards by video, and inside a long video by frame range with a causal halo of PS + PT - 2
pictures (picture k needs the raw features of pictures k-(PS+PT-2)..k: main.m:144-153, 183-199).
The partition is static and computed on the host; no collective sits on the data path.  The only
communication is an optional final gather of maps / camera motion to rank 0 (torch.distributed:
NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def matlab_round(x):
    return int(np.floor(x + 0.5)) if x >= 0 else -int(np.floor(-x + 0.5))


def halo_frames(fps, fast=False):
    """PS + PT - 2 with PS = round(0.3 fps), PT = round(7 PS) (main.m:119,129); fast version: PS = cvRound(0.3 fps),
    PT = cvRound(0.7 fps) (salmat.cpp:18-19, round half to even)."""
    if fast:
        import numpy as np
        return max(1, int(np.rint(fps * 0.3))) + max(1, int(np.rint(fps * 0.7))) - 2
    ps = matlab_round(fps * 0.3)
    return ps + matlab_round(ps * 7.0) - 2


def plan_videos(lengths, world):
    """Longest-first bin packing of whole clips onto `world` ranks.  Returns one list of clip indices per rank."""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i); load[r] += lengths[i]
    return out


def plan_frame_ranges(nframes, world, halo):
    """Contiguous frame ranges of one clip, one per rank.  Each entry: first/count = owned pictures,
    seek = picture to cds_seek() to, context = pictures fed through cds_clip_context before `first`."""
    base, extra = divmod(nframes, world)
    out, first = [], 0
    for r in range(world):
        count = base + (1 if r < extra else 0)
        seek = max(0, first - halo)
        out.append({"rank": r, "first": first, "count": count, "seek": seek, "context": first - seek})
        first += count
    return out


def process_frame_range(clip, plan, hdr, tok, tok_base, ctu_bytes):
    """Run one rank's entry of plan_frame_ranges on a SaliencyClip: seek, halo as context, then the owned range."""
    s, f, n = plan["seek"], plan["first"], plan["count"]
    if n == 0:
        return np.zeros((0, clip.h, clip.w), np.uint8 if clip.out_u8 else np.float32), np.zeros((0, 2), np.int32)
    clip.seek(s)
    if plan["context"]:
        clip.context(s, hdr[s:f], tok, tok_base[s:f + 1], ctu_bytes[s:f])
    return clip.process(f, hdr[f:f + n], tok, tok_base[f:f + n + 1], ctu_bytes[f:f + n])


def gather_to_rank0(local, dist=None, group=None):
    """Concatenate per-rank arrays (first axis, rank order) on rank 0; other ranks get None.
    `dist` is torch.distributed (initialised by the caller); without it the array is returned as is."""
    if dist is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    import torch
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    pad = np.zeros((mx,) + local.shape[1:], local.dtype)
    pad[:local.shape[0]] = local
    t = torch.from_numpy(pad).to(dev)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    if rank != 0:
        return None
    return np.concatenate([p.cpu().numpy()[:c] for p, c in zip(parts, counts)], axis=0)


def bind_to_gpu_numa(gpu_index):
    """Pin this process to the CPU cores local to `gpu_index` (from `nvidia-smi topo -m`), BEFORE pinned host buffers are
    allocated: first-touch then places them on the GPU's NUMA node, which keeps the device->host map copies of several
    ranks from funnelling through one socket.  Returns the core list, or None when the topology cannot be read."""
    import os
    import re
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=20).stdout
    except (OSError, subprocess.TimeoutExpired):
        return None
    out = re.sub(r"\x1b\[[0-9;]*m", "", out)
    header = None
    for line in out.splitlines():
        cols = re.split(r"\t+|\s{2,}", line.strip())
        if header is None and any("CPU Affinity" in c for c in cols):
            header = [c.strip() for c in cols]
            continue
        if header is not None and cols and cols[0] == f"GPU{gpu_index}":
            # the row has one more leading cell (its own name) than the header has GPU/NIC columns
            for cell in cols[1:]:
                if re.fullmatch(r"[0-9,\-]+", cell) and ("-" in cell or "," in cell):
                    cpus = set()
                    for part in cell.split(","):
                        lo, _, hi = part.partition("-")
                        cpus.update(range(int(lo), int(hi or lo) + 1))
                    avail = os.sched_getaffinity(0)
                    cpus &= avail
                    if cpus:
                        os.sched_setaffinity(0, cpus)
                        return sorted(cpus)
                    return None
    return None

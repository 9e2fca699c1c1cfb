"""Frame sharding across the GPUs of one box: every (frame, instance) problem is
independent, so frames are split in contiguous blocks, the model bank is replicated, and the only
collective is one all_gather of the fixed-stride outputs (mirrors evaluator.py:240-249)."""
import torch
import torch.distributed as dist


def frame_range(n_frames, rank, world):
    """Contiguous block of frames for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balanced_assignment(instances_per_frame, world):
    """Greedy bin-pack of frames by instance count -> list of frame-id lists per rank."""
    order = sorted(range(len(instances_per_frame)), key=lambda f: -instances_per_frame[f])
    loads, bins = [0] * world, [[] for _ in range(world)]
    for f in order:
        r = min(range(world), key=lambda i: loads[i])
        bins[r].append(f)
        loads[r] += instances_per_frame[f]
    return [sorted(b) for b in bins]


def gather_outputs(local, n_frames_total, group=None):
    """all_gather a per-frame output tensor [n_local, ...] sharded by frame_range -> [n_frames_total, ...].
    Shards are padded to the largest block so one fixed-size collective suffices (NCCL on GPU, gloo on CPU)."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [frame_range(n_frames_total, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)

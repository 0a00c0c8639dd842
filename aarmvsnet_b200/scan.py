"""Partitioning a scan over the GPUs of one box.

A scan is a list of reference views, each an independent forward (eval.py:106-113); the depth-plane recurrence itself
cannot be split.  Rank r of G processes a contiguous slice of the reference views on its own GPU and the only
communication is ONE gather of the [H,W] depth + confidence maps at the end (NCCL on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_views(num_views, world_size, rank):
    """Contiguous balanced slice of range(num_views) owned by `rank` (first num_views % world_size ranks get one more)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(num_views, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def shard_units(views_per_scan, world_size, rank):
    """Several scans in one job (the reference's eval.py walks a test list of scans, eval.py:56-62): the work units are the
    (scan, reference view) pairs of ALL scans in order, cut into contiguous balanced slices, so the ceiling of one scan
    (49 views on 8 GPUs: the slowest rank has 7, a 7.0x speed-up) does not repeat per scan -- two scans give ceil(98 / 8) = 13
    units on the slowest rank instead of 14.  Contiguous slices keep a rank inside one scan as long as possible (per-scan
    feature cache).  Returns this rank's list of (scan index, view index)."""
    units = [(s, v) for s, n in enumerate(views_per_scan) for v in range(n)]
    return [units[i] for i in shard_views(len(units), world_size, rank)]


def max_share(num_views, world_size):
    return -(-num_views // world_size)


def gather_maps(local_maps, num_views, group=None):
    """local_maps: [n_local, C, H, W] tensor (this rank's views in shard order; C = depth, confidence, ...).
    Returns [num_views, C, H, W] on every rank.  One all_gather of equally sized (padded) shares."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = shard_views(num_views, world, rank)
    if local_maps.shape[0] != len(mine):
        raise ValueError(f"rank {rank} owns {len(mine)} views but passed {local_maps.shape[0]} maps")
    if world == 1:
        return local_maps
    share = max_share(num_views, world)
    padded = local_maps.new_zeros((share,) + tuple(local_maps.shape[1:]))
    padded[: len(mine)] = local_maps
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([parts[r][: len(shard_views(num_views, world, r))] for r in range(world)], 0)


def run_scan(forward_view, num_views, group=None):
    """forward_view(view_index) -> [C, H, W] maps of that reference view (e.g. a closure over EMVSNet and a loader).
    Runs this rank's share and gathers the whole scan."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    maps = [forward_view(v) for v in shard_views(num_views, world, rank)]
    return gather_maps(torch.stack(maps), num_views, group)
